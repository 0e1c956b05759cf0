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

    def quantize(self, tensors, scales, outs=None):
        res = [torch.from_numpy(oracle.quantize_with_scale(t.float().numpy(), s.numpy())).view(torch.float8_e4m3fn)
               for t, s in zip(tensors, scales)]
        if outs is not None:
            for o, r in zip(outs, res):
                o.view(torch.uint8).copy_(r.view(torch.uint8))
            return list(outs)
        return res

    def attend(self, q8, k8, v, sq, sk, sv, sm_scale, p_mode, out_dtype, out=None, return_lse=True):
        """v: e4m3 with head scales sv, or (sv None: the "16bit" mode) the 16-bit tensor itself."""
        self.calls.append(("attend", tuple(q8.shape), tuple(k8.shape)))
        if sv is None:
            v_bytes, sv_np = None, None
            q = torch.from_numpy(oracle.dequantize(q8.view(torch.uint8).numpy(), sq.numpy())).double()
            k = torch.from_numpy(oracle.dequantize(k8.view(torch.uint8).numpy(), sk.numpy())).double()
            s = (q @ k.transpose(-1, -2)) * sm_scale
            o, lse = torch.softmax(s, dim=-1) @ v.double(), torch.logsumexp(s, dim=-1)
        else:
            o, lse = oracle.attention_block_ref(q8.view(torch.uint8).numpy(), k8.view(torch.uint8).numpy(),
                                                v.view(torch.uint8).numpy(), sq.numpy(), sk.numpy(), sv.numpy(),
                                                sm_scale=sm_scale)
        o = o.to(out_dtype)
        if out is not None:
            out.copy_(o)
            o = out
        return (o, lse.float()) if return_lse else o

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
