"""CPU-only checks of the drop-in boundary: the library loads, exports every
symbol include/toyfhe_b200.h declares, and its host-only helpers (ring
construction) agree with the oracle.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

import toyfhe_b200 as T
from oracle import toyfhe_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    T.build_library()
    return T.load_library()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "toyfhe_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(tfb_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations found"
    assert sorted(T.ABI_SYMBOLS) == declared
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"


def test_prime_chain_matches_reference_constructor(lib, q8, psi8):
    qs, psis = T.prime_chain(2 ** 14, [60] * 8)
    assert qs == q8 and psis == psi8
    # unsorted request: generation is in ascending-logq order, result in request order (crt.jl:283-291)
    req = [60, 40, 40, 60, 50]
    assert T.prime_chain(2 ** 13, req) == tuple(O.prime_chain(2 ** 13, req))
    # the mnist example's ring (infer.jl:97-105)
    N = 2 ** 13
    q0 = O.nextprime(2 ** 60 + 1, 2 * N)
    ps = O.nextprime(q0 + 2 * N, 2 * N)
    qs7, _ = T.prime_chain(N, [60, 40, 40, 40, 40, 40, 60])
    assert qs7[0] == q0 and qs7[-1] == ps


def test_minimal_root(lib):
    assert T.minimal_primitive_root(97, 8) == 33          # docs/src/man/background/rlwe.md:183-187
    assert T.minimal_primitive_root(65537, 4096) == O.minimal_primitive_root(65537, 4096)


def test_ndigits(lib, q8):
    assert T.ndigits(q8, 2) == 241
    assert T.ndigits([97], 1) == 7


def test_error_reporting_without_gpu(lib):
    h = ctypes.c_void_p()
    q = (ctypes.c_uint64 * 1)(96)
    psi = (ctypes.c_uint64 * 1)(33)
    rc = lib.tfb_ctx_create(0, 4, 1, q, psi, ctypes.byref(h))
    assert rc == 1 and b"1 (mod 2N)" in lib.tfb_last_error()
    q[0] = 97
    psi[0] = 5
    rc = lib.tfb_ctx_create(0, 4, 1, q, psi, ctypes.byref(h))
    assert rc == 1 and b"primitive" in lib.tfb_last_error()
    rc = lib.tfb_prime_chain(16, None, 0, None, None)
    assert rc == 1


def test_header_is_plain_c(tmp_path):
    """The drop-in boundary must bind from C (and therefore from Julia's ccall, cgo, JNI ...): the header compiles as C99 with
    no C++ or CUDA types, and a C translation unit can take the address of every declared entry point."""
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = os.path.join(root, "include", "toyfhe_b200.h")
    names = sorted(set(re.findall(r"\b(tfb_[a-z0-9_]+)\s*\(", open(hdr).read())))
    assert len(names) > 60
    src = tmp_path / "use.c"
    src.write_text('#include "toyfhe_b200.h"\nvoid* table[] = {\n' + "".join(f"    (void*){n},\n" for n in names) + "};\n")
    out = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-Wno-pedantic", "-I", os.path.join(root, "include"), "-c", str(src), "-o", str(tmp_path / "use.o")],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stderr


def test_c_host_links_and_calls_the_library(tmp_path, q8, psi8):
    """A C host (what a cgo / JNI / ccall binding amounts to) links against libtoyfhe_b200.so and runs the host-only entry
    points: same prime chain as the reference's NegacyclicRing(N, logqs) constructor (crt.jl:282-295)."""
    import subprocess
    from toyfhe_b200 import LIB_PATH
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "host.c"
    src.write_text(r'''
#include <stdio.h>
#include <inttypes.h>
#include "toyfhe_b200.h"
int main(void) {
    int32_t logqs[8] = {60, 60, 60, 60, 60, 60, 60, 60};
    uint64_t q[8], psi[8];
    int rc = tfb_prime_chain(16384, logqs, 8, q, psi);
    if (rc) { printf("error %d: %s\n", rc, tfb_last_error()); return 1; }
    printf("%d\n", tfb_version());
    for (int i = 0; i < 8; i++) printf("%" PRIu64 " %" PRIu64 "\n", q[i], psi[i]);
    return tfb_prime_chain(16384, logqs, 8, 0, psi) == 0;   /* a null output is an error code, not a crash */
}
''')
    exe = tmp_path / "host"
    libdir = os.path.dirname(LIB_PATH)
    out = subprocess.run(["gcc", "-std=c99", "-I", os.path.join(root, "include"), str(src), "-o", str(exe), "-L", libdir, "-ltoyfhe_b200",
                          f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True)
    assert run.returncode == 0, run.stdout + run.stderr
    lines = run.stdout.split("\n")
    got = [tuple(int(v) for v in ln.split()) for ln in lines[1:9]]
    assert got == list(zip(q8, psi8))
