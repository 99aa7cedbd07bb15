"""CPU: the DETR and LXMERT oracle restatements reproduce the committed outputs of the UNMODIFIED reference
generators (tests/golden/detr_tiny.npz, lxmert_tiny.npz made by oracle/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import detr_oracle as do, lxmert_oracle as lo
from util import rel_err


def _sd(g):
    return {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}


@pytest.mark.parametrize("norm", [True, False])
@pytest.mark.parametrize("s10", [True, False])
def test_detr_oracle_golden(golden_dir, norm, s10):
    g = np.load(os.path.join(golden_dir, "detr_tiny.npz"))
    r = do.generate_ours(_sd(g), do.DETR_TINY, torch.from_numpy(g["src"]), torch.from_numpy(g["pos"]), torch.from_numpy(g["tq"]),
                         normalize_self_attention=norm, apply_self_in_rule_10=s10)
    assert rel_err(r, g[f"R.n{int(norm)}s{int(s10)}"]) < 1e-5


@pytest.mark.parametrize("norm", [True, False])
@pytest.mark.parametrize("s10", [True, False])
def test_lxmert_oracle_golden(golden_dir, norm, s10):
    g = np.load(os.path.join(golden_dir, "lxmert_tiny.npz"))
    rtt, rti, _ = lo.generate_ours(_sd(g), lo.LXMERT_TINY, torch.from_numpy(g["ids"]), torch.from_numpy(g["feats"]),
                                   torch.from_numpy(g["boxes"]), normalize_self_attention=norm, apply_self_in_rule_10=s10)
    key = f"n{int(norm)}s{int(s10)}"
    assert rel_err(rtt, g["Rtt." + key]) < 1e-5 and rel_err(rti, g["Rti." + key]) < 1e-5
