"""ctypes access to the CPU oracle (oracle/libtrre_oracle.so) and to the real reference binaries
(oracle/_ref/trre, oracle/_ref/trre_dft).  Test infrastructure only."""
import ctypes, os, subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "libtrre_oracle.so")
REF_NFT = os.path.join(ROOT, "oracle", "_ref", "trre")
REF_DFT = os.path.join(ROOT, "oracle", "_ref", "trre_dft")

NFT, DFT = 0, 1
SCAN, MATCH = 0, 1

_lib = None


def build_oracle():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True,
                   stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            build_oracle()
        L = ctypes.CDLL(ORACLE_SO)
        L.trre_oracle_run.restype = ctypes.c_int
        L.trre_oracle_run.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_char_p, ctypes.c_size_t,
                                      ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t),
                                      ctypes.c_char_p, ctypes.c_size_t]
        L.trre_oracle_free.argtypes = [ctypes.c_void_p]
        L.trre_oracle_last_nstates.restype = ctypes.c_size_t
        L.trre_oracle_last_dstates.restype = ctypes.c_size_t
        L.trre_oracle_synth_text.restype = ctypes.c_int
        L.trre_oracle_synth_text.argtypes = [ctypes.c_uint64, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_char_p,
                                             ctypes.c_double, ctypes.c_int]
        _lib = L
    return _lib


def oracle_run(expr: bytes, data: bytes, flavor=NFT, mode=SCAN, all_=False):
    """-> (rc, stdout bytes, stderr text)"""
    L = lib()
    out = ctypes.c_void_p()
    out_n = ctypes.c_size_t()
    err = ctypes.create_string_buffer(512)
    rc = L.trre_oracle_run(expr, flavor, mode, int(all_), data, len(data),
                           ctypes.byref(out), ctypes.byref(out_n), err, len(err))
    res = ctypes.string_at(out, out_n.value) if out.value else b""
    if out.value:
        L.trre_oracle_free(out)
    return rc, res, err.value.decode("latin1")


def oracle_synth_text(seed, n, words=None, p_hit=0.0, lowercase_only=False) -> bytearray:
    """The synthetic corpus from the oracle library's own generator - what bench.py's reference arm
    uses, so that it never maps the product's library."""
    buf = bytearray(n)
    if n:
        arr = (ctypes.c_ubyte * n).from_buffer(buf)
        w = None if not words else "\n".join(words).encode("latin1")
        rc = lib().trre_oracle_synth_text(seed, ctypes.addressof(arr), n, w, float(p_hit), int(lowercase_only))
        del arr
        assert rc == 0
    return buf


def have_ref():
    return os.access(REF_NFT, os.X_OK) and os.access(REF_DFT, os.X_OK)


def ref_run(expr: bytes, data: bytes, flavor=NFT, mode=SCAN, all_=False, timeout=60):
    """Run the real reference binary. -> (rc, stdout, stderr)"""
    argv = [REF_NFT if flavor == NFT else REF_DFT]
    if mode == MATCH:
        argv.append("-m")
    if all_:
        argv.append("-a")
    argv.append(expr)
    p = subprocess.run(argv, input=data, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=timeout)
    return p.returncode, p.stdout, p.stderr.decode("latin1")


def bounded_run(expr: bytes, data: bytes, flavor=NFT, mode=SCAN, all_=False, cpu_s=5, mem_mb=1024, out_mb=16):
    """The reference binary (or, without it, the oracle) in a CHILD process under hard limits: CPU seconds, address space,
    and the size of stdout (a file).  Generators and catastrophic backtracking are legal inputs of `trre -ma` and can
    print gigabytes or run for hours; a test must never do that inside the pytest process (it took two GPU boxes down).
    -> (rc, stdout) or None when a limit was hit."""
    import resource, sys, tempfile

    def limits():
        resource.setrlimit(resource.RLIMIT_CPU, (cpu_s, cpu_s))
        resource.setrlimit(resource.RLIMIT_AS, (mem_mb << 20, mem_mb << 20))
        resource.setrlimit(resource.RLIMIT_FSIZE, (out_mb << 20, out_mb << 20))

    if have_ref():
        argv = [REF_NFT if flavor == NFT else REF_DFT] + (["-m"] if mode == MATCH else []) + (["-a"] if all_ else []) + [expr]
    else:
        code = ("import sys; sys.path.insert(0, %r); from oracle_util import oracle_run; "
                "rc, out, _ = oracle_run(%r, sys.stdin.buffer.read(), flavor=%d, mode=%d, all_=%r); "
                "sys.stdout.buffer.write(out); sys.exit(rc)" % (os.path.dirname(os.path.abspath(__file__)), expr, flavor, mode, bool(all_)))
        argv = [sys.executable, "-c", code]
    with tempfile.TemporaryFile() as fo:
        try:
            p = subprocess.run(argv, input=data, stdout=fo, stderr=subprocess.DEVNULL, preexec_fn=limits, timeout=4 * cpu_s + 10)
        except subprocess.TimeoutExpired:
            return None
        if p.returncode not in (0, 1):
            return None                                    # killed by a limit (SIGXCPU, SIGXFSZ, SIGSEGV on a failed malloc ...)
        fo.seek(0)
        out = fo.read()
    if len(out) >= (out_mb << 20) - 4096:
        return None
    return p.returncode, out
