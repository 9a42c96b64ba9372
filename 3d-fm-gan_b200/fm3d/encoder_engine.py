"""Encoder inference on the implicit-GEMM engine (placeholder until the fused plan lands:
returning None makes the callers run their differentiable PyTorch composition on the GPU)."""


def run_resnet(model, x):
    return None


def run_psp(model, x):
    return None
