"""Host-side logic that needs no GPU: the nnet mirror's state_dict ABI, the C-ABI library's exports, and the
'no CPU fallback' contract."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_model(**kw):
    from nnet.model.onl_tfm_enc_1dcnn_enc_linear_non_autoreg_pos_enc_l2norm import OnlineTransformerDADiarization
    args = dict(n_speakers=4, in_size=345, n_units=256, n_heads=4, enc_n_layers=4, dec_n_layers=2, dropout=0.1,
                has_mask=True, max_seqlen=500, dec_dim_feedforward=2048)
    args.update(kw)
    return OnlineTransformerDADiarization(**args)


def test_state_dict_abi_matches_reference():
    """Keys, shapes and dtypes equal those of the real reference model (dumped by tests/golden/make_golden.py)."""
    want = {}
    with open(os.path.join(ROOT, "tests", "golden", "fs_state_dict_abi.txt")) as f:
        for line in f:
            k, shape, dt = re.match(r"(\S+) (\(.*\)) (\S+)", line.strip()).groups()
            want[k] = (eval(shape), dt)
    sd = make_model().state_dict()
    got = {k: (tuple(v.shape), str(v.dtype).replace("torch.", "")) for k, v in sd.items()}
    assert got == want
    assert sum(p.numel() for p in make_model().parameters()) == 9_974_450


def test_same_seed_same_init_as_reference_construction_order():
    """Two constructions under one seed agree, and the oracle's random_state_dict loads strictly."""
    from oracle import fs_eend_oracle as O
    torch.manual_seed(0)
    a = make_model().state_dict()
    torch.manual_seed(0)
    b = make_model().state_dict()
    assert all(torch.equal(a[k], b[k]) for k in a)
    make_model().load_state_dict(O.random_state_dict(0), strict=True)


def test_library_exports_every_declared_symbol(built_lib):
    from fseend_b200 import native
    hdr = open(os.path.join(ROOT, "include", "fseend_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(fseend_[a-z0-9_]+)\s*\(", hdr)))
    assert declared == sorted(native.EXPORTS)
    L = ctypes.CDLL(built_lib)
    for name in declared:
        assert hasattr(L, name), name
    assert L.fseend_version() == 100


def test_no_cpu_fallback(built_lib):
    """Without a GPU the product path must fail loudly (never route through the oracle or torch CPU ops)."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from fseend_b200.native import FseendError
    from oracle import fs_eend_oracle as O
    m = make_model().eval()
    src, lens = O.synthetic_features(1, 32)
    with pytest.raises((RuntimeError, FseendError)):
        m.test(src, lens, 4)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "fs-eend_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, fn), encoding="utf-8").read()
                assert "oracle" not in txt.replace("oracle/", "").lower() or fn == "README.md", (dp, fn)


def make_stream_model():
    from nnet.model.streaming_tfm_enc_1dcnn_enc_linear_non_autoreg_pos_enc_l2norm import StreamingTransformerEDADiarization
    return StreamingTransformerEDADiarization(in_size=345, n_units=256, n_heads=4, enc_n_layers=4, dec_n_layers=2,
                                              dropout=0.1, has_mask=True, max_seqlen=500, dec_dim_feedforward=2048)


def test_streaming_state_dict_abi_matches_reference():
    want = {}
    with open(os.path.join(ROOT, "tests", "golden", "fs_stream_state_dict_abi.txt")) as f:
        for line in f:
            k, shape, dt = re.match(r"(\S+) (\(.*\)) (\S+)", line.strip()).groups()
            want[k] = (eval(shape), dt)
    got = {k: (tuple(v.shape), str(v.dtype).replace("torch.", "")) for k, v in make_stream_model().state_dict().items()}
    assert got == want


def test_copy_params_from_masked_to_streaming_covers_every_tensor():
    """Every streaming tensor has a masked-model source of the same shape (reference copy_params.py:7-62), the
    dead masked parameters (dec.encoder*, norm12) are the only ones left uncopied."""
    from nnet.model.streaming_tfm_enc_1dcnn_enc_linear_non_autoreg_pos_enc_l2norm import streaming_to_masked_key
    from nnet.utils.copy_params import copy_params_from_masked_to_streaming
    from oracle import fs_eend_oracle as O
    masked, stream = make_model(), make_stream_model()
    masked.load_state_dict(O.random_state_dict(3), strict=True)
    copy_params_from_masked_to_streaming(masked, stream)
    msd, ssd = masked.state_dict(), stream.state_dict()
    used = set()
    for k, v in ssd.items():
        mk = streaming_to_masked_key(k)
        used.add(mk)
        assert torch.equal(v, msd[mk]), k
    unused = {k for k in msd if k not in used}
    assert all(k.startswith(("dec.encoder", "dec.encoder_norm")) or ".norm12." in k for k in unused), unused


def test_bench_flop_model_matches_survey_figures():
    """28.33 GFLOP per sequence at S=6 (21.24 at S=4) causal-exact, counting `convert` as written
    (2*T*S*512*256).  bench.py counts the split-weight form actually executed (2*T*256*256), i.e. is conservative."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    for S, survey in ((6, 28.33e9), (4, 21.24e9)):
        as_written = 2.0 * 500 * S * 512 * 256
        executed = 2.0 * 500 * 256 * 256
        assert abs(b.total_flops(1, 500, S) + as_written - executed - survey) < 0.01e9
    # attention core only: 4*hd*T(T+1)/2 per (head, sequence): 32.1 MFLOP
    assert abs(b.algorithmic_flops("enc.attn_causal", 1, 500, 6) / 4 - 32.064e6) < 1e3
