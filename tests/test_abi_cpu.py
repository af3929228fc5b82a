"""CPU-only checks: the C-ABI library loads and exports exactly what include/interactron_b200.h
declares; the product path refuses to run without a GPU; config / theta surfaces."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    txt = open(os.path.join(ROOT, "include", "interactron_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(itn_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from interactron_b200 import _lib
    lib = _lib.load()
    declared = _header_functions()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(_lib.SIGNATURES) == declared, "ctypes prototypes must mirror the header one to one"
    assert lib.itn_version().decode().endswith("sm_100a")


def test_argument_errors_are_reported_not_thrown():
    """Error contract: negative return + message, no exception, no compute without a GPU."""
    import ctypes as C
    from interactron_b200 import _lib
    lib = _lib.load()
    rc = lib.itn_layernorm_fwd(None, None, None, None, None, None, None, 10, 256, 1, 0, 1e-5, None)
    assert rc == -1 and b"null" in lib.itn_last_error()
    d = _lib.GemmDesc()
    assert lib.itn_gemm_tf32(C.byref(d), None) == -1
    assert b"M,N,K" in lib.itn_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    import interactron_b200 as ib
    from interactron_b200._lib import ItnError
    from interactron_b200.ops import CudaOps
    with pytest.raises(ItnError):
        CudaOps()
    m = ib.build_model(ib.default_config("interactron_random", weights="synthetic").MODEL).eval()
    from interactron_b200.synthetic import synthetic_episode
    with pytest.raises(RuntimeError, match="no CPU path"):
        m.predict(synthetic_episode(0))
    with pytest.raises(RuntimeError, match="no CPU path"):
        m.forward(synthetic_episode(0))


def test_fast_weight_enumeration():
    """157 tensors / 14,798,296 elements, reference order, in_proj_* and backbone excluded
    (SURVEY.md appendix A)."""
    from interactron_b200 import modules as M
    det = M.DetectorHolder(1235)
    items = M.fast_weight_items(det)
    names = [n for n, _ in items]
    assert len(items) == 157 and sum(p.numel() for _, p in items) == 14_798_296
    assert not any("in_proj" in n or "backbone" in n for n in names)
    assert names[0] == "transformer.encoder.layers.0.self_attn.out_proj.weight"
    assert names[60] == "transformer.decoder.layers.0.self_attn.out_proj.weight"
    assert names[144:148] == ["transformer.decoder.norm.weight", "transformer.decoder.norm.bias",
                              "class_embed.weight", "class_embed.bias"]
    assert names[154:] == ["query_embed.weight", "input_proj.weight", "input_proj.bias"]


def test_config_surface_matches_reference_yaml():
    import interactron_b200 as ib
    from oracle import reference_harness as rh
    c = ib.default_config("interactron")
    assert c.MODEL.ADAPTIVE_LR == 0.001 and c.MODEL.PREDICT_ACTIONS == 1 and c.MODEL.SET_COST_BBOX == 5
    assert isinstance(c.MODEL.SET_COST_BBOX, int) and c.MODEL.BLOCK_SIZE == 2060
    if not rh.reference_available():
        pytest.skip("reference checkout not present")
    for name in ("interactron", "interactron_random", "single_frame_baseline", "multi_frame_baseline"):
        mine = ib.default_config(name).MODEL.dictionarize()
        ref = rh.reference_config(name).MODEL.dictionarize()
        assert mine == ref, name
        loaded = ib.get_config(os.path.join(rh.REFERENCE_ROOT, "configs", name + ".yaml"))
        assert loaded.dictionarize() == rh.reference_config(name).dictionarize()


def test_state_dict_layout_matches_reference():
    import interactron_b200 as ib
    from oracle import reference_harness as rh
    if not rh.reference_available():
        pytest.skip("reference checkout not present")
    for name in ("interactron_random", "interactron"):
        m = ib.build_model(ib.default_config(name, weights="synthetic").MODEL)
        ref = rh.build_reference_model(name, m.state_dict())
        a, b = m.state_dict(), ref.state_dict()
        assert list(a.keys()) == list(b.keys())
        assert all(a[k].shape == b[k].shape for k in a)


def test_target_labels_are_range_checked():
    """A label outside [0, classes) must raise (torch's gather / cross_entropy do in the reference) instead of
    becoming an out-of-bounds read in the matcher-cost / criterion kernels."""
    import pytest
    import torch
    from interactron_b200.criterion import _pack_targets
    ok = [{"labels": torch.tensor([1, 1235]), "boxes": torch.rand(2, 4)}, {"labels": torch.zeros(0, dtype=torch.long),
                                                                          "boxes": torch.zeros(0, 4)}]
    labels, boxes, off, sizes, offs = _pack_targets(ok, torch.device("cpu"), num_logits=1236)
    assert labels.tolist() == [1, 1235] and offs == [0, 2, 2]
    for bad in (1236, -1):
        with pytest.raises(IndexError):
            _pack_targets([{"labels": torch.tensor([3, bad]), "boxes": torch.rand(2, 4)}], torch.device("cpu"),
                          num_logits=1236)
