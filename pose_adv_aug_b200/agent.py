"""On-GPU sampling of the agent's augmentation choices.

The reference turns the agent's logits into probabilities with softmax, copies them to the host and calls
`np.random.choice(K, 1, p=probs[j])` per sample, scale then rotation (joint-train-pose-s-r-agent.py:252-271 and
:344-363) -- a D2H sync in the middle of config 3's critical path.  Here the uniforms are drawn from numpy's global
RandomState BEFORE the forward result is needed (same stream consumption: one uniform per `choice`, interleaved
scale/rotation per sample), copied H2D asynchronously, and `hgk_softmax_sample` does softmax + inverse-CDF search
on the device.  With the same `np.random.seed` the indices equal the reference's (the uniform falls within one fp32
rounding of a CDF step with probability ~1e-7 per draw).
"""
import numpy as np
import torch

from ._lib import get_lib, HGKError


def draw_uniforms(n, rng=None):
    """[n,2] fp64 uniforms in the order the reference consumes them (sample-major, scale then rotation)."""
    r = np.random if rng is None else rng
    return torch.from_numpy(np.asarray(r.random_sample(2 * n), dtype=np.float64).reshape(n, 2))


def softmax_sample(logits, u):
    """logits [N,K] CUDA fp32, u [N] fp64 uniforms (host or device) -> (probs [N,K] fp32, index [N] int64), on device."""
    if not isinstance(logits, torch.Tensor) or not logits.is_cuda:
        raise HGKError("agent.softmax_sample runs on CUDA tensors only (no CPU fallback)")
    lib = get_lib()
    l = logits.detach().contiguous().float()
    N, K = l.shape
    ud = u.to(device=l.device, dtype=torch.float64, non_blocking=True).contiguous()
    if ud.numel() != N:
        raise ValueError("need one uniform per row")
    probs = torch.empty_like(l)
    index = torch.empty(N, device=l.device, dtype=torch.int64)
    lib.check(lib.softmax_sample(l.data_ptr(), N, K, ud.data_ptr(), probs.data_ptr(), index.data_ptr(),
                                 torch.cuda.current_stream(l.device).cuda_stream), "hgk_softmax_sample")
    return probs, index


def sample_scale_rotation(pred_scale_distri, pred_rotation_distri, uniforms=None, rng=None):
    """The sampling block of joint-train-pose-s-r-agent.py:252-271: returns
    (scale_probs, rotation_probs, scale_index, rotation_index), all CUDA tensors (no host round trip)."""
    n = pred_scale_distri.shape[0]
    u = draw_uniforms(n, rng) if uniforms is None else uniforms
    if u.device.type == "cpu":
        u = u.pin_memory() if torch.cuda.is_available() else u
    ps, si = softmax_sample(pred_scale_distri, u[:, 0])
    pr, ri = softmax_sample(pred_rotation_distri, u[:, 1])
    return ps, pr, si, ri
