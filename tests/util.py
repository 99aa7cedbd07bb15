import numpy as np
import torch


def rel_err(x, ref, base=None):
    """max|x-ref| / max|base| (base defaults to ref): the metric of SURVEY.md §8c."""
    x = torch.as_tensor(np.asarray(x.detach().cpu() if isinstance(x, torch.Tensor) else x), dtype=torch.float64)
    ref = torch.as_tensor(np.asarray(ref.detach().cpu() if isinstance(ref, torch.Tensor) else ref), dtype=torch.float64)
    base = ref if base is None else torch.as_tensor(np.asarray(base), dtype=torch.float64)
    denom = base.abs().max().item()
    if denom == 0:
        return (x - ref).abs().max().item()
    return (x - ref).abs().max().item() / denom


def text_rel_err(R, R_ref):
    """R_text is I + small: measure the error against the off-identity part so the diagonal 1s do not mask it."""
    R_ref = torch.as_tensor(np.asarray(R_ref), dtype=torch.float64)
    eye = torch.eye(R_ref.shape[-1], dtype=torch.float64)
    return rel_err(R, R_ref, base=(R_ref - eye))


TOL = 1e-4   # north_star: outputs match the reference PyTorch path within 1e-4 relative (fp32)
