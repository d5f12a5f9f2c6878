"""Drop-in for the two torchac functions the reference calls (compress.py:136, decompress.py:93), same names and arguments:

    encode_float_cdf(cdf_float, sym, needs_normalization=True, check_input_bounds=False) -> bytes
    decode_float_cdf(cdf_float, byte_stream, needs_normalization=True)                   -> int16 tensor

cdf_float is pn_kit.pmf_to_cdf's output [..., Lp]; the whole tensor is ONE stream, as in torchac.  The reference hands over CPU
tensors (`.cpu()` at compress.py:134-135); they are moved to the GPU, the coder runs there (csrc/entropy.cu) and the reference's
return types come back.  pcc_b200.codec uses ops.pmf_to_cdf_u16 / range_encode directly (one stream per cloud, no host hop).
torchac itself is absent from this image: its published algorithm is restated (parity unpinned; the CPU oracle restates the same algorithm).
"""
import torch

from . import ops


def _prep(cdf_float, needs_normalization):
    if not needs_normalization:
        raise NotImplementedError("pcc_b200 torchac shim: only needs_normalization=True (the reference's call) is supported")
    if cdf_float.dim() < 2:
        raise ValueError("cdf_float must have shape [..., Lp]")
    cdf = cdf_float if cdf_float.is_cuda else cdf_float.cuda()
    Lp = cdf.shape[-1]
    return ops.cdf_to_u16(cdf.float().reshape(-1, Lp)).view(1, -1, Lp), Lp


def encode_float_cdf(cdf_float, sym, needs_normalization=True, check_input_bounds=False):
    cdf, Lp = _prep(cdf_float, needs_normalization)
    if check_input_bounds:
        if float(cdf_float.min()) < 0 or float(cdf_float.max()) > 1:
            raise ValueError("cdf_float outside [0, 1]")
        if int(sym.min()) < 0 or int(sym.max()) > Lp - 2:
            raise ValueError("sym outside [0, Lp - 2]")
    if sym.dtype != torch.int16:
        raise ValueError("sym must be an int16 tensor")
    if tuple(sym.shape) != tuple(cdf_float.shape[:-1]):
        raise ValueError("sym and cdf_float disagree on the leading dimensions")
    data, nbytes = ops.range_encode(cdf, sym.reshape(1, -1))
    n = int(nbytes[0])
    ops.check_stream_sizes([n], data.shape[1])
    return bytes(data[0, :n].cpu().numpy().tobytes())


def decode_float_cdf(cdf_float, byte_stream, needs_normalization=True):
    cdf, _ = _prep(cdf_float, needs_normalization)
    buf = torch.frombuffer(bytearray(byte_stream), dtype=torch.uint8) if len(byte_stream) else torch.zeros(1, dtype=torch.uint8)
    n = torch.tensor([len(byte_stream)], dtype=torch.int32)
    sym = ops.range_decode(cdf, buf.view(1, -1), n)
    out = sym.view(tuple(cdf_float.shape[:-1]))
    return out if cdf_float.is_cuda else out.cpu()
