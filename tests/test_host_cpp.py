"""The C++ mirror of the reference interfaces (rocksplicator_b200/host/): tests/cpp/host_tests.cpp restates the
reference's own gtest cases (see the header of that file).  CPU part: helpers + the replication protocol over
a counting DbWrapper; GPU part: the same topologies with GpuDB below the DbWrapper seam."""
import os
import subprocess

import pytest


@pytest.fixture(scope="module")
def host_tests():
    from rocksplicator_b200 import build
    _, exe = build.build_host()
    return exe


def _run(exe, mode, timeout):
    p = subprocess.run([exe, mode], capture_output=True, text=True, timeout=timeout)
    print(p.stdout[-4000:])
    print(p.stderr[-2000:])
    assert p.returncode == 0, p.stdout[-3000:]
    assert " 0 failures" in p.stdout


def test_host_helpers_and_protocol_cpu(host_tests):
    _run(host_tests, "cpu", 300)


@pytest.mark.gpu
def test_host_gpu_backed(host_tests):
    if os.environ.get("RSP_TEST_EMUL_LIB"):  # tests/test_emul_cpu.py: the same binary linked against the emulation
        host_tests = os.path.join(os.path.dirname(os.environ["RSP_TEST_EMUL_LIB"]), "host_tests_emul")
    _run(host_tests, "gpu-only", 600)


def test_stager_stress(tmp_path):
    """csrc/stager.h alone (std only): 64 callers, two classes, sync + async completions, every result checked; once
    more under ThreadSanitizer when the toolchain has it"""
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpp", "stager_stress.cpp")
    cxx = os.environ.get("CXX", "g++")
    exe = str(tmp_path / "stager_stress")
    subprocess.check_call([cxx, "-std=c++17", "-O2", "-pthread", src, "-o", exe])
    p = subprocess.run([exe, "64", "800"], capture_output=True, text=True, timeout=300)
    print(p.stdout, p.stderr)
    assert p.returncode == 0 and "wrong 0" in p.stdout
    exe_t = str(tmp_path / "stager_stress_tsan")
    if subprocess.call([cxx, "-std=c++17", "-O1", "-g", "-fsanitize=thread", "-pthread", src, "-o", exe_t],
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) == 0:
        p = subprocess.run([exe_t, "16", "300"], capture_output=True, text=True, timeout=600)
        print(p.stdout, p.stderr[-3000:])
        assert p.returncode == 0 and "wrong 0" in p.stdout and "ThreadSanitizer" not in p.stderr


def test_arena_allocator(tmp_path):
    """csrc/arena.h over malloc: 60 000 random allocations / releases, contents and invariants checked (no overlap, slabs
    tiled exactly, free neighbours coalesced, a whole slab served again after everything was released)"""
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpp", "arena_test.cpp")
    cxx = os.environ.get("CXX", "g++")
    exe = str(tmp_path / "arena_test")
    flags = ["-std=c++17", "-O1", "-g", "-pthread"]
    if subprocess.call([cxx] + flags + ["-fsanitize=address", src, "-o", exe], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) != 0:
        subprocess.check_call([cxx] + flags + [src, "-o", exe])
    p = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(p.stdout, p.stderr[-2000:])
    assert p.returncode == 0 and "bad 0" in p.stdout
