"""mtrl_b200: B200-native (sm_100a) implementation of the multi-task SAC update path of
reginald-mclean/mtrl, behind that project's own Python API (mtrl.rl / mtrl.nn / mtrl.config / replay
buffer).  The compute is hand-written CUDA reached through the C-ABI in include/mtrl_b200.h."""

__version__ = "0.1.0"
