"""The thread-per-read seeding code of the product (bioseqdb_b200/csrc/seed_thread.cuh) compiled for the HOST and run on a CPU copy of
the index against the oracle's mem_collect_intv: identical interval lists and identical bwt_extend counts for every read the code takes
(tests/seed_thread_check.cpp).  Runs without a GPU; the same code runs on the device in kernel `seed_thread`."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "_build", "seed_thread_check")


@pytest.fixture(scope="module")
def checker(oracle):
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-ftrivial-auto-var-init=pattern", "-Wno-unknown-pragmas", "-o", EXE, os.path.join(ROOT, "tests", "seed_thread_check.cpp"),
                           "-L" + os.path.join(ROOT, "oracle"), "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle")])
    return EXE


# ref_len n_reads read_len sub indel repeat_copies K seed wide
@pytest.mark.parametrize("args", [
    "200000 4000 150 0.01 0.001 0 8 1 0",      # unique reference
    "200000 4000 150 0.01 0.001 12 8 2 0",     # repeat families + a low-complexity stretch
    "300000 3000 150 0.03 0.005 40 9 3 0",     # noisy reads, many copies
    "100000 2000 150 0.01 0.001 10 6 4 1",     # 64-bit rows (texts of 2^32 symbols and more use them)
    "60000 2000 250 0.05 0.01 60 5 13 0",      # long noisy reads, shallow table: real extensions everywhere
])
def test_thread_seeding_equals_oracle(checker, args):
    p = subprocess.run([checker] + args.split(), capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    f = p.stdout.split()
    reads, mism, fallback = int(f[1]), int(f[3]), int(f[5])
    assert mism == 0 and int(f[8]) == int(f[10]) and int(f[8]) > 0
    assert fallback < reads        # the thread code really took reads
