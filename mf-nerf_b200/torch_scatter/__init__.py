"""Drop-in for the one torch_scatter entry point the reference uses: `segment_csr(src, indptr)` (sum reduction) in
RayMarcher.backward (models/custom_functions.py:4,108-110).  Backed by mfn_segment_sum in libmfnerf_b200.so."""
import torch

from mfnerf_b200._lib import call, ptr, stream_ptr

__all__ = ["segment_csr"]


def segment_csr(src, indptr, out=None, reduce="sum"):
    if reduce not in ("sum", "add"):
        raise NotImplementedError(f"segment_csr reduce={reduce!r}")
    if not src.is_cuda:
        raise RuntimeError("src must be a CUDA tensor (no CPU fallback)")
    if indptr.dim() != 1:
        raise NotImplementedError("segment_csr: only 1-D indptr")
    s = src.float().contiguous()
    width = int(s[0].numel()) if s.shape[0] > 0 else int(torch.tensor(s.shape[1:]).prod().item()) if s.dim() > 1 else 1
    n_seg = indptr.shape[0] - 1
    res = torch.empty((n_seg,) + tuple(s.shape[1:]), dtype=torch.float32, device=s.device) if out is None else out
    with torch.cuda.device(s.device):
        call("mfn_segment_sum", ptr(s), ptr(indptr.to(torch.int64).contiguous()), n_seg, width, ptr(res), stream_ptr(s.device))
    return res.to(src.dtype)
