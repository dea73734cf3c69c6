"""Fused PointTransformer local aggregation (placeholder until the fused kernels land):
falls back to nothing — raising keeps the contract that there is no silent fallback."""


def pt_attention(layer, p, x_q, x_k, x_v, idx):
    raise NotImplementedError("fused pt_attention not built yet")
