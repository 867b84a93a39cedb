"""The switches table of DESIGN.md and the getenv() calls of the sources must name the same variables (a switch that is
documented but gone, or present but undocumented, is a stale document)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _read(*parts):
    with open(os.path.join(ROOT, *parts), encoding="utf-8") as f:
        return f.read()


def _source_switches():
    names = set()
    csrc = os.path.join(ROOT, "mercer_research_b200", "csrc")
    for fn in os.listdir(csrc):
        if fn.endswith((".cu", ".cuh")):
            names |= set(re.findall(r'getenv\("(RCN_[A-Z0-9_]+)"', _read("mercer_research_b200", "csrc", fn)))
    pkg = os.path.join(ROOT, "mercer_research_b200")
    for fn in [os.path.join(pkg, f) for f in os.listdir(pkg) if f.endswith(".py")] + [os.path.join(ROOT, "bench.py")]:
        with open(fn, encoding="utf-8") as f:
            names |= set(re.findall(r'environ(?:\.get\(|\[)\s*"(RCN_[A-Z0-9_]+)"', f.read()))
    return names


def test_every_switch_in_the_sources_is_documented():
    documented = set(re.findall(r"RCN_[A-Z0-9_]+", _read("DESIGN.md") + _read("README.md") + _read("INTEGRATION.md")))
    missing = sorted(n for n in _source_switches() if n not in documented and n != "RCN_SERVE_LOG")
    assert not missing, f"undocumented environment switches: {missing}"


def test_every_documented_switch_exists():
    design = _read("DESIGN.md")
    start = design.index("### Build-free switches")
    table = design[start:design.index("\n## ", start)]
    documented = set(re.findall(r"`(RCN_(?:CUDA|BENCH)_[A-Z0-9_]+)", table))
    assert documented, "switch table not found"
    gone = sorted(documented - _source_switches())
    assert not gone, f"documented switches that no source reads: {gone}"
