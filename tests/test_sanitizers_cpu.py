"""Host-side code under sanitizers (SURVEY.md 5 "race detection / sanitizers"):
the schedule planner (csrc/pbd_plan.cpp + pbd_tileplan.cpp + pbd_placement.cpp, multithreaded tile colouring and placement) is built
host-only with -fsanitize=address,undefined and -fsanitize=thread and driven by tools/plan_sanitize.cpp
through every order mode; the reference-side adapter (integration/CudaStepper.cpp) must compile against
the reference's own header with -Wall -Wextra -Wpedantic -Werror."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = [os.path.join(ROOT, "tools", "plan_sanitize.cpp"),
       os.path.join(ROOT, "cs121-softbodysim_b200", "csrc", "pbd_plan.cpp"),
       os.path.join(ROOT, "cs121-softbodysim_b200", "csrc", "pbd_tileplan.cpp"),
       os.path.join(ROOT, "cs121-softbodysim_b200", "csrc", "pbd_placement.cpp")]


@pytest.mark.parametrize("name,flags,n", [("asan_ubsan", "-fsanitize=address,undefined", "12"), ("tsan", "-fsanitize=thread", "20")])
def test_planner_under_sanitizers(name, flags, n, tmp_path):
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("g++ not available")
    exe = str(tmp_path / f"plan_{name}")
    cmd = [gxx, "-std=c++17", "-O1", "-g", flags, "-fno-sanitize-recover=all", "-fno-omit-frame-pointer", "-pthread", *SRC, "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1:halt_on_error=1",
               TSAN_OPTIONS="halt_on_error=1")
    # n = 20 (48k tets) is above the planner's multi-threading threshold (200k constraints), so TSan sees its worker threads
    run = subprocess.run([exe, n], capture_output=True, text=True, timeout=900, env=env)
    assert run.returncode == 0 and run.stdout.strip().endswith("OK"), (run.stdout[-1500:], run.stderr[-3000:])
    assert "ERROR: AddressSanitizer" not in run.stderr and "runtime error" not in run.stderr and "WARNING: ThreadSanitizer" not in run.stderr


def test_reference_adapter_compiles_without_warnings():
    ref = "/root/reference/CProgram/include/PBDServer.h"
    if not os.path.exists(ref):
        pytest.skip("the reference tree is not on this machine (the adapter is built where it is)")
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "integration"), "check"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    assert "warning" not in (r.stdout + r.stderr).lower()
