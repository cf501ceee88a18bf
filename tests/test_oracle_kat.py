"""The oracle is pinned against the reference's own known-answer tests (oracle/kat_tests.cpp)."""

import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_oracle_known_answer_tests():
    subprocess.run(["make", "-s", "-C", str(ROOT / "oracle")], check=True)
    r = subprocess.run([str(ROOT / "oracle" / "kat_tests")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failed" in r.stdout
