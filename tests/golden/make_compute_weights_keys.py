"""Generates tests/golden/compute_weights_keys.json: the keys of the metrics dictionary the reference's
`MTSAC.compute_weights` returns (/root/reference/mtrl/rl/algorithms/mtsac.py:1085-1170), parsed from its source in the
build container (the function itself needs jax).  Run once: `python tests/golden/make_compute_weights_keys.py`."""
import json
import os
import re

SRC = "/root/reference/mtrl/rl/algorithms/mtsac.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "compute_weights_keys.json")

src = open(SRC).read()
blk = src[src.index("        return self, {\n            # Existing metrics"):]
blk = blk[: blk.index("\n        }\n")]
keys = sorted(set(re.findall(r'^\s+"([a-z_]+)":', blk, re.M)))
json.dump(keys, open(OUT, "w"), indent=0)
print(len(keys), "keys")
