"""bench.py contract on CPU: the reference arm prints ONE JSON line with the keys the driver reads (metric, unit,
value, impl, cpu_baseline, e2e with zero copies), and the files README/DESIGN cite under profiles/ exist."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "end-to-end PPO env-steps/sec" and d["unit"] == "env-steps/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


def test_cited_profile_files_exist():
    cited = set()
    for doc in ("README.md", "DESIGN.md", os.path.join("profiles", "README.md")):
        text = open(os.path.join(ROOT, doc)).read()
        cited |= set(re.findall(r"profiles/(r01_[A-Za-z0-9_]+\.(?:json|csv|txt))", text))
        if doc.startswith("profiles"):
            cited |= set(re.findall(r"`(r01_[A-Za-z0-9_]+\.(?:json|csv|txt))`", text))
    assert cited, "no profile files cited"
    missing = sorted(f for f in cited if not os.path.exists(os.path.join(ROOT, "profiles", f)))
    assert not missing, missing
