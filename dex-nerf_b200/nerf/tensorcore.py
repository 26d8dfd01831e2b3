"""Host side of the tcgen05 tensor-core MLP path (dex-nerf_b200/csrc/mlp_tc.cu)."""


def supported(model, prog):
    return False


def query(model, prog, ro, rd, viewdirs, z, rf):
    raise NotImplementedError
