"""Builds the TEST-ONLY mock engine (mock_bpmsm.c: the C ABI of include/bpmsm.h on the oracle's CPU arithmetic) and a copy of the C++ host
mirror linked against it.  Test infrastructure: called by tests/conftest.py and __graft_entry__.build(), never by the package."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
MOCK = os.path.join(HERE, "libmockbpmsm.so")
HOST = os.path.join(HERE, "libbulletproofs_host_mock.so")


def _stale(target, sources):
    return not os.path.exists(target) or any(os.path.getmtime(s) > os.path.getmtime(target) for s in sources)


def _run(cmd):
    print("+", " ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)


def build_mock_engine(force=False):
    odir, hdir, csrc = os.path.join(ROOT, "oracle"), os.path.join(ROOT, "bulletproofs_b200", "host"), os.path.join(ROOT, "bulletproofs_b200", "csrc")
    srcs = [os.path.join(HERE, f) for f in ("mock_bpmsm.c", "mock_stubs.c")]
    deps = srcs + [os.path.join(odir, f) for f in ("ge.h", "sc.h", "fe51.h", "hashes.h", "liboracle.so")] + [os.path.join(ROOT, "include", "bpmsm.h")]
    if force or _stale(MOCK, deps):
        _run(["gcc", "-O2", "-fPIC", "-shared", "-o", MOCK] + srcs + ["-L" + odir, "-loracle", "-Wl,-rpath,$ORIGIN/../../oracle"])
    hsrcs = [os.path.join(hdir, f) for f in ("bulletproofs.cpp", "r1cs.cpp", "mpc.cpp")]
    hdeps = [os.path.join(hdir, f) for f in os.listdir(hdir)] + [os.path.join(csrc, f) for f in ("sc.cuh", "merlin.cuh", "fe.cuh")] + [MOCK]
    if force or _stale(HOST, hdeps):
        _run(["g++", "-O2", "-std=c++17", "-pthread", "-shared", "-fPIC", "-o", HOST] + hsrcs + ["-L" + HERE, "-lmockbpmsm", "-Wl,-rpath,$ORIGIN"])
    return MOCK, HOST


if __name__ == "__main__":
    build_mock_engine(force=True)
