"""Autograd wrapper of the CUDA HGTConv forward (backward kernels: see csrc/edge_bwd.cu when present)."""


def hgt_conv_autograd(module, node_inp, node_type, edge_index, edge_type, edge_time):
    raise NotImplementedError(
        "pyhgt_b200.HGTConv: the backward pass is not implemented yet; call forward under "
        "torch.no_grad() / module.eval() with requires_grad disabled")
