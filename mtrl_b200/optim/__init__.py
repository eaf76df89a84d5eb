"""The reference's mtrl.optim holds optax-style gradient-surgery transforms (pcgrad, gradnorm,
cagrad, dummy) with the protocol init(params)->state / update(updates, state, params, **extra).
They need per-task gradients and are a SURVEY 8(f) "next" row; the plain Adam + clip chain of the
hot path is fused in csrc/sac_kernels.cuh (adam_kernel)."""
