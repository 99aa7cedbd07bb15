"""CPU: the C-ABI library builds (cross-compile), loads, and exports every symbol include/mmx.h declares.
No compute calls - there is no GPU here."""
import os
import re

import pytest


def _declared_symbols():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "include", "mmx.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mmx_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(built_lib):
    import mmx_b200
    decl = _declared_symbols()
    assert len(decl) >= 20
    for name in decl:
        assert hasattr(built_lib, name), f"{name} declared in include/mmx.h but not exported by libmmx.so"
    assert sorted(mmx_b200.exported_symbols()) == decl, "ctypes binding and header disagree"
    assert built_lib.mmx_version() == 1


def test_no_cpu_fallback_without_gpu(built_lib):
    import torch
    import mmx_b200
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from oracle import clip_oracle as co
    with pytest.raises(mmx_b200.MmxError):
        mmx_b200.ClipEngine(mmx_b200.ClipConfig(*co.TINY.ref_args()), co.init_state_dict(co.TINY), max_batch=1)
    with pytest.raises(mmx_b200.MmxError):
        mmx_b200.avg_heads(torch.zeros(2, 3, 3), torch.zeros(2, 3, 3))


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under the product package may import, load or exec it (Python) or
    include / link it (C++ / CUDA).  Checked on the syntax tree for Python and on #include lines for native code."""
    import ast
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "transformer-mm-explainability_b200")
    checked = 0
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            path = os.path.join(dirpath, f)
            if f.endswith(".py"):
                tree = ast.parse(open(path).read(), path)
                for node in ast.walk(tree):
                    names = []
                    if isinstance(node, ast.Import):
                        names = [a.name for a in node.names]
                    elif isinstance(node, ast.ImportFrom):
                        names = [node.module or ""]
                    elif isinstance(node, ast.Call) and getattr(node.func, "attr", getattr(node.func, "id", "")) in (
                            "import_module", "__import__", "spec_from_file_location"):
                        names = [a.value for a in node.args if isinstance(a, ast.Constant) and isinstance(a.value, str)]
                    for n in names:
                        assert n.split(".")[0] != "oracle" and "oracle/" not in n and "/oracle" not in n, (path, n)
                checked += 1
            elif f.endswith((".cu", ".cuh", ".h", ".cpp", ".c")):
                for line in open(path):
                    if line.lstrip().startswith("#include"):
                        assert "oracle" not in line, (path, line)
                checked += 1
    assert checked >= 15


def test_config_from_state_dict():
    import mmx_b200
    from oracle import clip_oracle as co
    for cfg in (co.TINY, co.SMALL):
        sd = co.init_state_dict(cfg)
        got = mmx_b200.ClipConfig.from_state_dict(sd)
        # build_model infers text heads as width // 64 (CLIP/clip/model.py:430); tiny configs use other head
        # counts, so compare everything except that field
        a, b = list(cfg.ref_args()), [got.embed_dim, got.image_resolution, got.vision_layers, got.vision_width,
                                      got.vision_patch_size, got.context_length, got.vocab_size,
                                      got.transformer_width, got.transformer_heads, got.transformer_layers]
        a[8] = b[8] = 0
        assert a == b
