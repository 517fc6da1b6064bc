"""bench.py's roofline numerators are the as-written FLOP counts of SURVEY.md 8(d): pinned here so that the reported TFLOP/s
cannot drift from the survey's figures (cfg 2: 1.302 GF/clip, cfg 3: 2.221, cfg 4: 2.256, cfg 5: 2.952; per cell-step cfg 2:
4.063 MF layer 0, 3.171 MF layer 1; cfg 5: 25.26 / 28.37 MF)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def test_flop_counts_match_survey():
    import bench
    want = {2: 1.302, 3: 2.221, 4: 2.256, 5: 2.952}
    for c, gf in want.items():
        assert abs(bench.flops_per_clip(bench.CONFIGS[c]) / 1e9 - gf) < 5e-4, c
    assert abs(bench.fcell(100 + 64, 64, 1, 2) / 1e6 - 4.063) < 5e-4
    assert abs(bench.fcell(64 + 64, 64, 1, 2) / 1e6 - 3.171) < 5e-4
    assert abs(bench.fcell(100 + 128, 128, 2, 3) / 1e6 - 25.26) < 5e-3
    assert abs(bench.fcell(128 + 128, 128, 2, 3) / 1e6 - 28.37) < 5e-3


def test_kernel_families_cover_the_whole_step():
    """the algorithmic FLOPs attributed to the kernels of the default (second-generation) path -- x pre-projection + recurrent
    forward, BPTT + dX, weight gradient -- add up to the fwd+bwd count of the step, minus the input gradient of layer 0, which
    the as-written count (3 x forward) contains but nobody needs (layer 0's dX is discarded)"""
    import bench
    cfg = bench.CONFIGS[2]
    fam = bench.kernel_families(cfg)
    path = ("xproj", "rnn_fwd", "rnn_bwd", "dx16", "dw_mm16")
    total = sum(sum(fam[k]) for k in path)
    step = bench.flops_per_clip(cfg) * cfg["B"]
    dx_layer0 = fam["xproj"][0]
    assert abs(total + dx_layer0 - step) < 1e-6 * step, (total, dx_layer0, step)
