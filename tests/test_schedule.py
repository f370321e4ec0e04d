"""Host logic on the CPU: the chunk schedule of an analysis call (sdft_b200/csrc/sdft_chunks.hpp), compiled with
g++ on its own and checked for its invariants over 200 000 random (cursor, n, m, chunk length) combinations."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_chunk_schedule_invariants(tmp_path):
    exe = str(tmp_path / "schedule_check")
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "sdft_b200", "csrc"),
                    os.path.join(ROOT, "tests", "drivers", "schedule_check.cpp"), "-o", exe], check=True)
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout
    assert "schedules ok" in res.stdout
