"""The deciding arithmetic of the quad scene's closest-hit scan (quasi_b200/csrc/scan_decide.hpp,
the code the CUDA kernel compiles) run on the host against the literal loop of the reference
shader: whenever the fast scan does not hand a ray to the literal loop, it must name the
literal winner. tests/scan_decide_check.cpp aims the rays at the hard places (edges of the
blocks' faces, the floor within t_min of a block, directions almost parallel to a face)."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def build(tmp_path, include_dir, *flags):
    exe = tmp_path / ("scan_decide_check" + "".join(f.replace("-D", "_").replace("=", "_") for f in flags))
    subprocess.run(["g++", "-std=c++20", "-O2", "-ffp-contract=off", "-Wall", "-Wextra", f"-I{include_dir}", *flags,
                    str(ROOT / "tests" / "scan_decide_check.cpp"), "-o", str(exe)], check=True)
    return exe


@pytest.mark.parametrize("form", ["quads", "boxes"])
def test_fast_scan_names_the_literal_winner(tmp_path, form):
    """Both forms of the scan: every quad on its own (the default) and blocks as slab tests."""
    exe = build(tmp_path, ROOT / "quasi_b200" / "csrc")
    out = subprocess.run([str(exe), "4000000", form], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-2000:]
    counts = dict(zip(out.stdout.split()[0::2], out.stdout.split()[1::2]))
    assert int(counts["mismatches"]) == 0 and int(counts["decided"]) > 2_000_000
    assert int(counts["literal"]) > 0          # the margins do fire on these rays


def test_margins_absorb_an_approximate_reciprocal(tmp_path):
    """On the device 1/x in the deciding arithmetic is one MUFU.RCP (relative error 2^-23) and
    a*b+c is one FFMA; on the host they are exact quotient and fmaf. With the reciprocal pushed
    two ulps either way and the FFMAs replaced by separately rounded operations the scan must
    still never name a quad the literal loop does not."""
    exe = build(tmp_path, ROOT / "quasi_b200" / "csrc", "-DQUASI_SCAN_CHECK_PERTURB")
    out = subprocess.run([str(exe), "4000000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-2000:]
    assert "mismatches 0" in out.stdout


def test_the_check_notices_a_missing_grazing_rule(tmp_path):
    """The speed-scaled edge margin of scan_box (a ray almost in the plane of the face it enters
    through may be counted as hitting a neighbouring face earlier: one wrong quad in 2e8
    uniformly aimed rays without it) has rays aimed at it too."""
    mutated = tmp_path / "mutated_grazing"
    mutated.mkdir()
    src = (ROOT / "quasi_b200" / "csrc" / "scan_decide.hpp").read_text()
    assert "edge < 0.0f" in src and "(maybe ? 0.0f : k_inf)" in src
    src = src.replace("edge < 0.0f", "outside < -b.delta").replace("(maybe ? 0.0f : k_inf)", "(maybe ? t_entry : k_inf)")
    (mutated / "scan_decide.hpp").write_text(src)
    (mutated / "scene_params.hpp").write_text((ROOT / "quasi_b200" / "csrc" / "scene_params.hpp").read_text())
    exe = build(tmp_path, mutated)
    out = subprocess.run([str(exe), "12000000"], capture_output=True, text=True)
    assert out.returncode == 1 and "MISMATCH" in out.stdout


def test_the_check_notices_a_missing_rule(tmp_path):
    """Without the rule that sends rays starting within t_min of a block's face to the literal
    loop, the scan names the wrong quad for some of them: the check must see that."""
    mutated = tmp_path / "mutated"
    mutated.mkdir()
    src = (ROOT / "quasi_b200" / "csrc" / "scan_decide.hpp").read_text()
    assert "(parallel || tunnels)" in src
    (mutated / "scan_decide.hpp").write_text(src.replace("(parallel || tunnels)", "(parallel)"))
    (mutated / "scene_params.hpp").write_text((ROOT / "quasi_b200" / "csrc" / "scene_params.hpp").read_text())
    exe = build(tmp_path, mutated)
    out = subprocess.run([str(exe), "3000000"], capture_output=True, text=True)
    assert out.returncode == 1 and "MISMATCH" in out.stdout


def test_the_check_notices_a_missing_per_ray_rule(tmp_path):
    """The shader's |denom| < 1e-8 rejection (plugin.mm:174) and its t < t_min rejection inside the
    band t' cannot resolve (plugin.mm:177) are taken once per ray, on the winner
    (scan_decide.hpp:needs_literal_scan / accept_winner), not once per quad. The check aims rays
    at both: grazing rays that start within 1e-8 of a quad's plane, and rays that meet a quad at
    t_min give or take 1e-5 of it. Without either rule the scan names quads the shader rejects."""
    csrc = ROOT / "quasi_b200" / "csrc"
    src = (csrc / "scan_decide.hpp").read_text()
    assert "|| s.best_t <= k_t_min_hi;" in src and "return std::fabs(denom) >= 1e-8f; }" in src
    for name, old, new in (("band", "|| s.best_t <= k_t_min_hi;", ";"),
                           ("denom", "return std::fabs(denom) >= 1e-8f; }", "return denom == denom; }")):
        mutated = tmp_path / f"mutated_{name}"
        mutated.mkdir()
        (mutated / "scan_decide.hpp").write_text(src.replace(old, new))
        (mutated / "scene_params.hpp").write_text((csrc / "scene_params.hpp").read_text())
        exe = build(tmp_path, mutated, f"-DMUTATION_{name}")
        out = subprocess.run([str(exe), "3000000", "quads"], capture_output=True, text=True)
        assert out.returncode == 1 and "MISMATCH" in out.stdout, name
