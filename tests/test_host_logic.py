"""CPU tests (-m "not gpu"): C-ABI surface, host-side logic, and that nothing computes without a GPU."""
import ctypes as C
import os
import re
import numpy as np
import pytest
import ifrit_v2_b200  # noqa: F401
from ifrit_v2_b200 import renderer as R, scenes, multigpu as MG

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_library_loads_and_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "ifrit_b200.h")).read()
    names = sorted(set(re.findall(r"\b(iftrcu[A-Za-z0-9]+)\s*\(", hdr)))
    assert len(names) >= 45
    lib = R.load_library()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ifrit_b200.h but not exported"


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(R.IfritError) as e:
        R.TileRasterRendererCuda(0)
    assert e.value.code == R.IFRIT_ERR_CUDA


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may touch oracle/."""
    pkg = os.path.join(ROOT, "ifrit-v2_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")) or f == "Makefile":
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
                assert "libifrit_oracle" not in src and "ifrit_oracle.h" not in src, f


def test_pod_struct_layouts_match_reference_headers():
    # engine/base/Structures.h:52-97 as compiled by g++ on x86-64
    assert C.sizeof(R.IfritSamplerT) == 28
    assert C.sizeof(R.IfritImageCreateInfo) == 28
    assert C.sizeof(R.IfritImageBlit) == 64
    assert C.sizeof(R.IfritBufferImageCopy) == 36
    assert C.sizeof(R.IfritColorAttachmentBlendState) == 20


def test_shader_registry_lookup():
    lib = R.load_library()
    assert lib.iftrcuLookupVertexShader(b"mvp") == R.VS_MVP
    assert lib.iftrcuLookupFragmentShader(b"textured") == R.FS_TEXTURED
    assert lib.iftrcuLookupFragmentShader(b"nope") == -1


def test_scenes_are_seeded_and_shaped():
    vb, idx = scenes.mesh_sphere(n=160)
    assert vb.shape == (25921, 12) and idx.shape == (51200 * 3,) and idx.max() == 25920
    a, _ = scenes.soup(ntri=1000, seed=42)
    b, _ = scenes.soup(ntri=1000, seed=42)
    assert np.array_equal(a, b) and a.shape == (3000, 8)
    t = scenes.texture_checker(size=64)
    assert t.shape == (64, 64, 4) and t.dtype == np.float32 and 0 <= t.min() and t.max() <= 1.0


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_owned_tile_bytes_matches_host_partition(world):
    for (w, h) in ((3840, 2160), (7680, 4320), (400, 230), (64, 64)):
        total = 0
        for rank in range(world):
            n = len(MG.owned_supertiles(w, h, rank, world))
            assert R.owned_tile_bytes(w, h, rank, world) == n * 64 * 64 * 16
            total += n
        assert total == ((w + 63) // 64) * ((h + 63) // 64)   # a partition: every super-tile owned once


def test_pack_unpack_host_roundtrip():
    rng = np.random.default_rng(0)
    frame = rng.random((230, 400, 4), dtype=np.float32)
    for world in (2, 3, 8):
        out = np.zeros_like(frame)
        for rank in range(world):
            MG.unpack_host(out, MG.pack_host(frame, rank, world), rank, world)
        assert np.array_equal(out, frame)


def test_bench_algorithmic_bytes_follow_survey_8d():
    """bench.py's roofline numerator: B = V*Sv + I*4 + P*(16+4) + Tex, Tex = min(1.33*covered, texels in chain)*16
    (SURVEY.md §8d); C2 figures quoted there: 1.2 MB + 0.6 MB + 165.9 MB + <= 89.5 MB."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    cfg = bench.CONFIGS["c2"]
    sc = dict(vb=np.zeros((25921, 12), np.float32), idx=np.zeros(153600, np.int32), tex=np.zeros((2048, 2048, 4), np.float32), mips=11)
    chain = sum(((2048 + (1 << l) - 1) >> l) ** 2 for l in range(12))
    ab = bench.algorithmic_bytes(cfg, sc, covered_pixels=4_691_104)        # C2: 1.33 x covered exceeds the chain
    assert ab["geometry"] == 25921 * 48 + 153600 * 4
    assert ab["tex"] == chain * 16 and abs(ab["tex"] - 89.5e6) < 0.1e6
    assert ab["tile"] == 3840 * 2160 * 20 + ab["tex"]
    assert ab["frame"] == ab["geometry"] + ab["tile"]
    few = bench.algorithmic_bytes(cfg, sc, covered_pixels=1_000_000)
    assert few["tex"] == int(1.33 * 1_000_000) * 16
    assert set(bench.CONFIGS) == {"c2", "c3", "c4", "c5"} and bench.CONFIGS["c3"]["msaa"] == 8
