"""Device versions of the tail of the reference's feature pipeline, FS-EEND/datasets/feature.py (same in LS-EEND):
``splice`` (:111-133) and ``subsample`` (:103-108) fused into one kernel (csrc/elementwise.cu, through the C ABI).
STFT + mel + log (librosa in the reference) are NOT part of this module — SURVEY §8f N3 is only started.
"""
import torch


def splice_subsample(Y, context_size=7, subsampling=10):
    """Equivalent of ``subsample(splice(Y, context_size), T, subsampling)[0]`` on the GPU.
    Y: (n_frames, n_featdim) float tensor / array -> CUDA fp32 (ceil(n_frames / subsampling), n_featdim * (2 ctx + 1))."""
    from fseend_b200.native import op_splice_subsample
    if not torch.cuda.is_available():
        raise RuntimeError("fseend_b200 runs on a CUDA sm_100 device only")
    y = torch.as_tensor(Y).detach().to(device="cuda", dtype=torch.float32).contiguous()
    return op_splice_subsample(y, context_size, subsampling)
