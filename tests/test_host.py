"""CPU tests of the product's host side: the C-ABI library loads and exports every symbol of
include/gaselect_b200.h, the host RNG / CV segmentation are bit-exact against the golden vectors
produced by the reference's own code, argument validation mirrors the reference constructors,
and -- with no GPU -- the hot path fails loudly instead of falling back to the CPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, cuda_device_count
from helpers import make_data, seg_digest


def _product_evaluator(spec, X, y, seed):
    from gaselect_b200 import evaluators as E
    spec = dict(spec)
    kind = spec.pop("kind")
    if kind == "pls":
        return E.PLSEvaluator(X, y, seed, **spec)
    if kind == "fit":
        return E.BICEvaluator(X, y, seed, **spec)
    return E.LMEvaluator(X, y, **spec)


def test_library_exports_every_declared_symbol(engine):
    from gaselect_b200 import _capi
    header = open(os.path.join(ROOT, "include", "gaselect_b200.h")).read()
    declared = set(re.findall(r"^(?:int|void)\s+(gsb_\w+)\s*\(", header, flags=re.M))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(engine, name), name + " is declared in include/gaselect_b200.h but not exported"
    assert declared == set(_capi.SIGNATURES), "ctypes signatures out of sync with the header"
    assert engine.gsb_abi_version() == 4


def test_struct_sizes_match_the_header(engine, tmp_path):
    """The header is plain C: compile it with gcc and compare the struct sizes with ctypes'."""
    import subprocess
    from gaselect_b200 import _capi
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "gaselect_b200.h"\nint main(void){printf("%zu %zu %zu\\n",'
                   'sizeof(gsb_config),sizeof(gsb_info),sizeof(gsb_timing));return 0;}\n')
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert sizes == [C.sizeof(_capi.Config), C.sizeof(_capi.Info), C.sizeof(_capi.Timing)]


def test_host_rng_known_answers(engine, golden):
    from gaselect_b200 import _capi
    g = golden["rng"]
    sv = _capi.seed_vector(golden["ga_seed"])
    assert sv[:4].tolist() == g["seedvec123_head"] and int(sv[623]) == g["seedvec123_tail"]
    assert int(np.bitwise_xor.reduce(sv)) == g["seedvec123_xor"]
    out = np.zeros(8, dtype=np.uint32)
    assert engine.gsb_rng_stream(sv, 8, out) == 0
    assert out.tolist() == g["vec_first8"]
    for seed, first in g["u32_first8"].items():          # RNG(uint32) is what seedvec() draws from
        assert _capi.seed_vector(int(seed))[:8].tolist() == first
    # literal SURVEY.md Appendix C.1 values
    assert _capi.seed_vector(123)[:4].tolist() == [2866969740, 3565251014, 4277846663, 460894080]
    out6 = np.zeros(6, dtype=np.uint32)
    engine.gsb_rng_stream(_capi.seed_vector(123), 6, out6)
    assert out6.tolist() == [2738463805, 2572576382, 2132272039, 2119152319, 2475546466, 737790329]
    long = np.zeros(3000, dtype=np.uint32)
    engine.gsb_rng_stream(sv, 3000, long)
    assert long[:8].tolist() == g["vec_first8"]


def test_host_rng_long_stream_matches_oracle(engine, oracle_lib):
    from gaselect_b200 import _capi
    for seed in (0, 1, 123, 4294967295):
        sv = _capi.seed_vector(seed)
        assert (sv == oracle_lib.make_seedvec(seed)).all()
        got = np.zeros(5000, dtype=np.uint32)
        engine.gsb_rng_stream(sv, 5000, got)
        assert (got == oracle_lib.rng_vec(sv, 5000)).all()


def test_host_shuffle(engine, golden):
    from gaselect_b200 import _capi
    sv = _capi.seed_vector(golden["ga_seed"])
    out = np.zeros(3 * 40, dtype=np.uint32)
    assert engine.gsb_shuffle_all(sv, 40, 3, out) == 0
    assert out.reshape(3, 40).tolist() == golden["shuffle40x3"]
    out = np.zeros(2 * 150, dtype=np.uint32)
    engine.gsb_shuffle_all(sv, 150, 2, out)
    got = [int(np.bitwise_xor.reduce(r * np.arange(1, 151, dtype=np.uint32))) for r in out.reshape(2, 150)]
    assert got == golden["shuffle150x2_xor"]


def _case_names():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")) as fh:
        return [c["name"] for c in json.load(fh)["cases"] if c["evaluator"]["kind"] != "lm"]


@pytest.mark.parametrize("name", _case_names())
def test_segmentation_bit_exact(engine, golden, name):
    """gsb_init_segmentation == PLSEvaluator/BICEvaluator::initSegmentation of the reference."""
    case = next(c for c in golden["cases"] if c["name"] == name)
    X, y = make_data(case["data"])
    with _product_evaluator(case["evaluator"], X, y, golden["ga_seed"]) as ev:
        seg = ev.getSegmentation()
        assert len(seg) == case["segmentation_count"]
        assert seg_digest(seg) == case["segmentation_digest"]
        if "segmentation" in case:
            assert [s.tolist() for s in seg] == case["segmentation"]
        inf = ev.info()
        assert inf["n"] == X.shape[0] and inf["p"] == X.shape[1]


def test_segmentation_shape_clamps(engine):
    from gaselect_b200 import evaluators as E, synth
    X, y = synth.closed_form()
    with E.PLSEvaluator(X, y, 123, numReplications=3, maxNComp=0, innerSegments=7, outerSegments=1) as ev:
        inf = ev.info()   # I_eff = 6, maxNComp -> 28
        assert (inf["inner_segments"], inf["max_ncomp"], inf["segmentation_vectors"]) == (6, 28, 42)
        assert inf["sdfact_effective"] == pytest.approx(1.0 / np.sqrt(6.0), rel=1e-15)
    with E.PLSEvaluator(X, y, 123, numReplications=2, maxNComp=5, innerSegments=4, outerSegments=3) as ev:
        inf = ev.info()
        assert (inf["inner_segments"], inf["outer_segments"], inf["max_ncomp"], inf["segmentation_vectors"]) == (4, 3, 5, 60)


def test_set_segmentation_round_trip(engine):
    from gaselect_b200 import evaluators as E, synth
    X, y = synth.closed_form()
    with E.PLSEvaluator(X, y, 123, numReplications=2, maxNComp=5, innerSegments=4, outerSegments=3) as a, \
            E.PLSEvaluator(X, y, 999, numReplications=2, maxNComp=5, innerSegments=4, outerSegments=3) as b:
        seg = a.getSegmentation()
        assert any((s != t).any() for s, t in zip(seg, b.getSegmentation()) if len(s) == len(t))
        b.setSegmentation(seg)
        assert all((s == t).all() for s, t in zip(seg, b.getSegmentation()))
        with pytest.raises(ValueError):
            b.setSegmentation(seg[:-2])


def test_invalid_arguments_mirror_the_reference_constructors(engine):
    from gaselect_b200 import evaluators as E, synth
    X, y = synth.closed_form()
    with pytest.raises(ValueError):      # src/PLSEvaluator.cpp:50-52
        E.PLSEvaluator(X, y, 1, testSetSize=1.0)
    with pytest.raises(ValueError):
        E.PLSEvaluator(X, y, 1, testSetSize=-0.1)
    with pytest.raises(ValueError):      # src/BICEvaluator.cpp:31-33
        E.BICEvaluator(X, y, 1, numSegments=1)
    with pytest.raises(ValueError):
        E.PLSEvaluator(X, y[:-1], 1)
    with pytest.raises(KeyError):
        E.LMEvaluator(X, y, statistic="nope")
    E.PLSEvaluator(X, y, 1, outerSegments=4, testSetSize=7.0).close()   # rdCV ignores testSetSize (:46-48)


def test_bad_population_is_rejected_before_any_cuda_call(engine):
    from gaselect_b200 import evaluators as E, synth
    if cuda_device_count() > 0:
        pytest.skip("needs the no-GPU environment to prove no CUDA call is involved")
    X, y = synth.closed_form()
    with E.LMEvaluator(X, y) as ev:
        with pytest.raises(E.EngineError) as err:   # prepare comes first and needs the device
            ev.evaluate([1, 2, 3])
        assert err.value.code == 3


def test_no_gpu_means_loud_failure_not_cpu_fallback(engine):
    from gaselect_b200 import evaluators as E, synth
    if cuda_device_count() > 0:
        pytest.skip("a CUDA device is present")
    X, y = synth.closed_form()
    for ev in (E.PLSEvaluator(X, y, 123, numReplications=1, innerSegments=3), E.BICEvaluator(X, y, 123), E.LMEvaluator(X, y)):
        with pytest.raises(E.EngineError) as err:
            ev.evaluate_population([[1, 2, 3]])
        assert err.value.code == 3  # GSB_ERR_CUDA
        with pytest.raises(E.EngineError):
            ev.simpls([1, 2, 3], 2)
        ev.close()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gaselect_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("oracle-independent", ""), os.path.join(dirpath, f)


def test_multi_device_deal_is_a_balanced_deterministic_partition(engine):
    """gsb_deal: the deal of a multi-device engine (pure host function): a partition with equal counts, balanced on the
    cost model, deterministic, and the same deal as the Python host of the one-process-per-GPU path (multigpu.deal)."""
    from gaselect_b200 import _capi, multigpu, synth
    subsets = synth.random_subsets(1000, 501, 0, 200, seed=11)
    _, offsets = _capi.to_csr(subsets)
    sizes = np.diff(offsets.astype(np.int64))
    for ndev in (1, 2, 3, 8):
        for kspace in (0, 1):
            dev = np.full(len(subsets), -1, dtype=np.int32)
            assert engine.gsb_deal(offsets, len(subsets), ndev, kspace, dev) == 0
            assert dev.min() >= 0 and dev.max() < ndev
            counts = np.bincount(dev, minlength=ndev)
            assert counts.max() <= -(-len(subsets) // ndev) and counts.sum() == len(subsets)
            cost = multigpu.kspace_cost(sizes) if kspace else multigpu.gram_cost(sizes)
            loads = np.array([cost[dev == d].sum() for d in range(ndev)])
            assert loads.max() - loads.min() <= 1.5 * cost.max()
            again = np.zeros_like(dev)
            engine.gsb_deal(offsets, len(subsets), ndev, kspace, again)
            assert (again == dev).all()
            shares = multigpu.deal(sizes, ndev, cost=cost, equal_counts=True)
            for d in range(ndev):
                assert np.flatnonzero(dev == d).tolist() == shares[d].tolist()
    assert engine.gsb_deal(offsets, len(subsets), 0, 0, dev) != 0 and engine.gsb_deal(offsets, len(subsets), 17, 0, dev) != 0


def test_multi_device_configuration_is_validated(engine):
    """A device list with a repeated or negative ordinal is an invalid argument; a valid list creates an engine without
    touching CUDA (host-side segmentation works; evaluation fails loudly without the devices)."""
    from gaselect_b200 import evaluators as E
    from gaselect_b200 import synth
    X, y = synth.nir_spectra(40, 12, seed=1)
    with pytest.raises(ValueError):
        E.PLSEvaluator(X, y, 1, numReplications=2, innerSegments=3, devices=[0, 0])
    with pytest.raises(ValueError):
        E.PLSEvaluator(X, y, 1, numReplications=2, innerSegments=3, devices=[0, -1])
    with E.PLSEvaluator(X, y, 1, numReplications=2, innerSegments=3, devices=[0, 1]) as multi:
        with E.PLSEvaluator(X, y, 1, numReplications=2, innerSegments=3) as single:
            assert seg_digest(multi.getSegmentation()) == seg_digest(single.getSegmentation())
        if cuda_device_count() < 2:
            with pytest.raises(E.EngineError):
                multi.evaluate_population([np.arange(5, dtype=np.uint32)])
