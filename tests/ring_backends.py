"""Oracle-backed stand-in for quantumattention_b200.parallel.NativeBackend, so the ring's HOST logic (scale exchange,
block rotation, merge order) can run on CPU under gloo.  Test infrastructure only."""
import numpy as np
import torch

import oracle


class OracleBackend:
    def __init__(self):
        self.calls = []

    def local_scales(self, tensors):
        return [torch.from_numpy(oracle.head_scales(t.float().numpy())) for t in tensors]

    def quantize(self, tensors, scales):
        return [torch.from_numpy(oracle.quantize_with_scale(t.float().numpy(), s.numpy())).view(torch.float8_e4m3fn)
                for t, s in zip(tensors, scales)]

    def attend(self, q8, k8, v8, sq, sk, sv, sm_scale, p_mode, out_dtype):
        self.calls.append("attend")
        o, lse = oracle.attention_block_ref(q8.view(torch.uint8).numpy(), k8.view(torch.uint8).numpy(),
                                            v8.view(torch.uint8).numpy(), sq.numpy(), sk.numpy(), sv.numpy(),
                                            sm_scale=sm_scale)
        return o.to(out_dtype), lse.float()

    def merge(self, o_acc, lse_acc, o_new, lse_new, first, out=None):
        self.calls.append("merge_first" if first else "merge")
        if first:
            o, lse = o_new.double(), lse_new.double()
        else:
            o, lse = oracle.merge_ref(o_acc.double(), lse_acc.double(), o_new.double(), lse_new.double())
        lse_acc.copy_(lse.float())
        if out is not None:
            out.copy_(o.to(out.dtype))
        else:
            o_acc.copy_(o.float())
