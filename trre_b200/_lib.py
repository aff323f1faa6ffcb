"""ctypes binding of include/trre_b200.h (the drop-in C ABI)."""
import ctypes
import os

FLAVOR_NFT, FLAVOR_DFT = 0, 1
EXEC_BYTEMAP, EXEC_TABLE, EXEC_BACKTRACK, EXEC_LAZY = 0, 1, 2, 3
FAST_NONE, FAST_BYTEMAP, FAST_MATCH, FAST_BYTEMAP_VAR, FAST_MATCH_VAR, FAST_INPLACE = 0, 1, 2, 3, 4, 5

OK, E_PATTERN, E_UNSUPPORTED, E_CUDA, E_NOSPACE, E_STACK, E_ARG, E_ASSERT = 0, 1, 2, 3, 4, 5, 7, 134

_HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path():
    # TRRE_B200_LIB: an alternative build of the same library (experiments with compile-time knobs)
    return os.environ.get("TRRE_B200_LIB") or os.path.join(_HERE, "lib", "libtrre_b200.so")


class TrreError(RuntimeError):
    """code = TRRE_E_* ; text = what the reference would have printed on stderr (for E_PATTERN)."""

    def __init__(self, code, text):
        super().__init__("trre_b200 error %d: %s" % (code, text.strip()))
        self.code = code
        self.text = text


class _Info(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint32) for n in (
        "abi_version", "flavor", "exec_kind", "nft_states", "dsa_states", "byte_classes",
        "max_depth", "length_preserving", "can_match_empty", "table_bytes", "fast_path", "gram_tables",
        "literal_words")]


_L = None


def _lib():
    global _L
    if _L is not None:
        return _L
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError("%s is missing: build it with `make` or __graft_entry__.build(); "
                          "the trre_b200 scan path has no Python/CPU fallback" % path)
    L = ctypes.CDLL(path)
    vp, sz, u8p = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p
    L.trre_compile.restype = ctypes.c_int
    L.trre_compile.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(vp), ctypes.c_char_p, sz]
    L.trre_prog_free.argtypes = [vp]
    L.trre_prog_free.restype = None
    L.trre_prog_get_info.restype = ctypes.c_int
    L.trre_prog_get_info.argtypes = [vp, ctypes.POINTER(_Info)]
    L.trre_prog_blob.restype = vp
    L.trre_prog_blob.argtypes = [vp, ctypes.POINTER(sz)]
    L.trre_prog_from_blob.restype = ctypes.c_int
    L.trre_prog_from_blob.argtypes = [vp, sz, ctypes.POINTER(vp), ctypes.c_char_p, sz]
    L.trre_lazy_build.restype = ctypes.c_int
    L.trre_lazy_build.argtypes = [vp, ctypes.POINTER(ctypes.c_uint32), sz, ctypes.c_uint32]
    L.trre_lazy_view.restype = ctypes.c_int
    L.trre_lazy_view.argtypes = [vp, ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(vp), ctypes.POINTER(vp),
                                 ctypes.POINTER(vp), ctypes.POINTER(sz)]
    L.trre_prog_set_mode.restype = ctypes.c_int
    L.trre_prog_set_mode.argtypes = [vp, ctypes.c_int, ctypes.c_int]
    L.trre_prog_bytemap.restype = ctypes.c_int
    L.trre_prog_bytemap.argtypes = [vp, ctypes.c_char_p]
    L.trre_scan.restype = ctypes.c_int
    L.trre_scan.argtypes = [vp, ctypes.c_int, u8p, sz, ctypes.POINTER(vp), ctypes.POINTER(sz)]
    L.trre_scan_multi.restype = ctypes.c_int
    L.trre_scan_multi.argtypes = [vp, ctypes.POINTER(ctypes.c_int), ctypes.c_int, u8p, sz, ctypes.POINTER(vp), ctypes.POINTER(sz)]
    L.trre_scan_into.restype = ctypes.c_int
    L.trre_scan_into.argtypes = [vp, ctypes.c_int, u8p, sz, u8p, sz, ctypes.POINTER(sz)]
    L.trre_scan_device.restype = ctypes.c_int
    L.trre_scan_device.argtypes = [vp, ctypes.c_int, u8p, sz, u8p, sz, ctypes.POINTER(sz), vp]
    L.trre_scan_status.restype = ctypes.c_int
    L.trre_scan_status.argtypes = [vp, ctypes.c_int, ctypes.POINTER(sz), vp]
    L.trre_scan_launches.restype = ctypes.c_int
    L.trre_scan_launches.argtypes = [vp, ctypes.c_int]
    L.trre_shard_lines.restype = ctypes.c_int
    L.trre_shard_lines.argtypes = [u8p, sz, ctypes.c_int, ctypes.POINTER(sz)]
    L.trre_match_host.restype = ctypes.c_int
    L.trre_match_host.argtypes = [vp, ctypes.c_int, ctypes.c_int, u8p, sz, ctypes.POINTER(vp),
                                  ctypes.POINTER(sz), ctypes.c_char_p, sz]
    L.trre_synth_text.restype = ctypes.c_int
    L.trre_synth_text.argtypes = [ctypes.c_uint64, u8p, sz, ctypes.c_char_p, ctypes.c_double, ctypes.c_int]
    L.trre_free.argtypes = [vp]
    L.trre_free.restype = None
    L.trre_last_error.restype = ctypes.c_char_p
    L.trre_version.restype = ctypes.c_char_p
    _L = L
    return L


def version():
    return _lib().trre_version().decode()


def _last_error():
    return _lib().trre_last_error().decode("latin1")


def _as_buffer(data):
    """bytes / bytearray / memoryview / numpy array -> (address, nbytes, keepalive)"""
    if isinstance(data, bytes):
        return ctypes.cast(ctypes.c_char_p(data), ctypes.c_void_p).value or 0, len(data), data
    mv = memoryview(data).cast("B")
    if mv.readonly:
        b = bytes(mv)
        return ctypes.cast(ctypes.c_char_p(b), ctypes.c_void_p).value or 0, len(b), b
    arr = (ctypes.c_ubyte * len(mv)).from_buffer(mv)
    return ctypes.addressof(arr), len(mv), (arr, mv)


class Program:
    """A compiled trre expression: flat NFT + deterministic scan automaton, resident per device."""

    def __init__(self, handle):
        self._h = ctypes.c_void_p(handle)
        info = _Info()
        rc = _lib().trre_prog_get_info(self._h, ctypes.byref(info))
        if rc:
            raise TrreError(rc, "trre_prog_get_info failed")
        self.info = {n: getattr(info, n) for n, _ in _Info._fields_}

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _L is not None:
            _L.trre_prog_free(h)

    # ---- host side ---------------------------------------------------------------------------
    @property
    def exec_kind(self):
        return self.info["exec_kind"]

    @property
    def flavor(self):
        return self.info["flavor"]

    def blob(self) -> bytes:
        n = ctypes.c_size_t()
        p = _lib().trre_prog_blob(self._h, ctypes.byref(n))
        return ctypes.string_at(p, n.value)

    def bytemap(self) -> bytes:
        buf = ctypes.create_string_buffer(256)
        rc = _lib().trre_prog_bytemap(self._h, buf)
        if rc:
            raise TrreError(rc, "program is not a byte map")
        return buf.raw

    def lazy_build(self, states, extra=0):
        """TRRE_EXEC_LAZY: build the rows of `states` (+ `extra` more, breadth first)."""
        arr = (ctypes.c_uint32 * len(states))(*states)
        rc = _lib().trre_lazy_build(self._h, arr, len(states), extra)
        if rc:
            raise TrreError(rc, _last_error())

    def lazy_view(self):
        """TRRE_EXEC_LAZY: (n_states, rows address, sinfo address, pool address, pool bytes) of the tables built so far."""
        n, rows, sinfo, pool, pb = ctypes.c_uint32(), ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_size_t()
        rc = _lib().trre_lazy_view(self._h, ctypes.byref(n), ctypes.byref(rows), ctypes.byref(sinfo), ctypes.byref(pool), ctypes.byref(pb))
        if rc:
            raise TrreError(rc, "not a lazily built program")
        return n.value, rows.value, sinfo.value, pool.value, pb.value

    def set_mode(self, match=False, all_=False):
        """The CLI's -m / -a: every later scan*() of this program runs that mode on the GPU (DFS executor)."""
        rc = _lib().trre_prog_set_mode(self._h, int(match), int(all_))
        if rc:
            raise TrreError(rc, _last_error())
        return self

    def match(self, data, all_=False, scan=False) -> bytes:
        """`-m` / `-ma` (scan=False) or scan-mode `-a` (scan=True) on the HOST: cross-check of set_mode() + scan()."""
        addr, n, keep = _as_buffer(data)
        out, out_n = ctypes.c_void_p(), ctypes.c_size_t()
        err = ctypes.create_string_buffer(256)
        rc = _lib().trre_match_host(self._h, int(scan), int(all_), addr, n, ctypes.byref(out),
                                    ctypes.byref(out_n), err, len(err))
        res = ctypes.string_at(out, out_n.value) if out.value else b""
        if out.value:
            _lib().trre_free(out)
        if rc:
            e = TrreError(rc, err.value.decode("latin1"))
            e.partial = res
            raise e
        return res

    # ---- the scan path (GPU) -----------------------------------------------------------------
    def scan(self, data, device=0) -> bytes:
        """`trre EXPR < data` (or trre_dft): host buffer in, host bytes out; H2D + kernels + D2H."""
        addr, n, keep = _as_buffer(data)
        out, out_n = ctypes.c_void_p(), ctypes.c_size_t()
        rc = _lib().trre_scan(self._h, device, addr, n, ctypes.byref(out), ctypes.byref(out_n))
        if rc:
            raise TrreError(rc, _last_error())
        res = ctypes.string_at(out, out_n.value) if out.value else b""
        _lib().trre_free(out)
        return res

    def scan_multi(self, data, devices) -> bytes:
        """The buffer cut at line boundaries, shard k on devices[k], outputs joined in order."""
        addr, n, keep = _as_buffer(data)
        devs = (ctypes.c_int * len(devices))(*devices)
        out, out_n = ctypes.c_void_p(), ctypes.c_size_t()
        rc = _lib().trre_scan_multi(self._h, devs, len(devices), addr, n, ctypes.byref(out), ctypes.byref(out_n))
        if rc:
            raise TrreError(rc, _last_error())
        res = ctypes.string_at(out, out_n.value) if out.value else b""
        _lib().trre_free(out)
        return res

    def scan_into(self, in_addr, n, out_addr, out_cap, device=0) -> int:
        """Host pointers (e.g. pinned torch tensors' data_ptr()); returns bytes written."""
        out_n = ctypes.c_size_t()
        rc = _lib().trre_scan_into(self._h, device, in_addr, n, out_addr, out_cap, ctypes.byref(out_n))
        if rc:
            e = TrreError(rc, _last_error())
            e.needed = out_n.value
            raise e
        return out_n.value

    def scan_device(self, d_in, n, d_out, out_cap, device=0, stream=0, sync=True):
        """Device pointers.  sync=True returns the output size; sync=False only enqueues."""
        out_n = ctypes.c_size_t()
        rc = _lib().trre_scan_device(self._h, device, d_in, n, d_out, out_cap,
                                     ctypes.byref(out_n) if sync else None, stream)
        if rc:
            e = TrreError(rc, _last_error())
            e.needed = out_n.value
            raise e
        return out_n.value if sync else None

    def scan_status(self, device=0, stream=0) -> int:
        out_n = ctypes.c_size_t()
        rc = _lib().trre_scan_status(self._h, device, ctypes.byref(out_n), stream)
        if rc:
            e = TrreError(rc, _last_error())
            e.needed = out_n.value
            raise e
        return out_n.value

    def launches(self, device=0) -> int:
        return _lib().trre_scan_launches(self._h, device)


def compile(expr, flavor=FLAVOR_NFT) -> Program:
    if isinstance(expr, str):
        expr = expr.encode("latin1")
    h = ctypes.c_void_p()
    err = ctypes.create_string_buffer(512)
    rc = _lib().trre_compile(expr, flavor, ctypes.byref(h), err, len(err))
    if rc:
        raise TrreError(rc, err.value.decode("latin1"))
    return Program(h.value)


def from_blob(blob: bytes) -> Program:
    h = ctypes.c_void_p()
    err = ctypes.create_string_buffer(256)
    rc = _lib().trre_prog_from_blob(blob, len(blob), ctypes.byref(h), err, len(err))
    if rc:
        raise TrreError(rc, err.value.decode("latin1"))
    return Program(h.value)


def shard_lines(data, parts):
    """Line-aligned cut points: list of parts+1 offsets."""
    addr, n, keep = _as_buffer(data)
    cuts = (ctypes.c_size_t * (parts + 1))()
    rc = _lib().trre_shard_lines(addr, n, parts, cuts)
    if rc:
        raise TrreError(rc, "trre_shard_lines failed")
    return list(cuts)


def synth_text(seed, n, words=None, p_hit=0.0, lowercase_only=False) -> bytearray:
    """Deterministic synthetic corpus: printable ASCII, lines U{40..120}, ends in \\n."""
    buf = bytearray(n)
    if n:
        arr = (ctypes.c_ubyte * n).from_buffer(buf)
        w = None if not words else "\n".join(words).encode("latin1")
        rc = _lib().trre_synth_text(seed, ctypes.addressof(arr), n, w, float(p_hit), int(lowercase_only))
        del arr
        if rc:
            raise TrreError(rc, "trre_synth_text failed")
    return buf
