"""Test-only interpreter of the compiled program blob (trre_b200/csrc/automaton.hpp layout).

It lets the CPU test-suite validate the HOST compiler (parser, NFT, both determinizers, analysis)
against the oracle without a GPU: it executes the very tables the CUDA kernels walk, with the same
per-attempt rule documented at the top of automaton.hpp.  It is never used by the product."""
import struct

DEAD = 0x00FFFFFF
E_MARK, E_DECIDED = 1 << 24, 1 << 25
S_HAS_FINAL, S_DECIDED, S_DIVERGE = 1, 2, 4
PF_LP, PF_EMPTY, PF_ACYCLIC, PF_NUL, PF_DIVERGE = 1, 2, 4, 8, 16

HDR = struct.Struct("<IIIIII I i IIII IIIIIIII 4I")
FIELDS = ("magic abi total_bytes flavor exec_kind flags nft_states nft_start dsa_states n_classes max_depth "
          "start_info off_classmap off_lut off_startrow off_trans off_sinfo off_pool pool_bytes off_nft").split()


# header words beyond FIELDS (automaton.hpp BlobHeader): index -> name
EXT = {28: "swar_n", 29: "swar0", 30: "swar1", 31: "swar2", 32: "swar3", 33: "gram_depth", 34: "gram_entries",
       35: "gram_len", 36: "gram_flags", 37: "off_gbits", 38: "off_gout", 39: "off_gpool", 40: "gpool_bytes",
       41: "lit_words", 42: "lit_len", 59: "vrun_classes", 23: "off_vent", 25: "off_vpool", 26: "vpool_bytes"}
EXT.update({43 + k: "lit%d" % k for k in range(16)})
EXT.update({60 + k: "lit_out%d" % k for k in range(4)})
EXT[64] = "gram_delta"
EXT[65] = "vrun_swar_n"
EXT.update({66 + k: "vrun_swar%d" % k for k in range(4)})
PF_MATCHABLE, PF_GRAM = 64, 512


class Diverged(Exception):
    pass


class Blob:
    def __init__(self, raw: bytes):
        vals = HDR.unpack_from(raw, 0)
        self.h = dict(zip(FIELDS, vals))
        h = self.h
        assert h["magic"] == 0x32425254 and h["total_bytes"] == len(raw)
        self.raw = raw
        self.classmap = raw[h["off_classmap"]:h["off_classmap"] + 256]
        self.lut = raw[h["off_lut"]:h["off_lut"] + 256]
        self.ncls = h["n_classes"]
        ns = h["dsa_states"]
        self.trans = struct.unpack_from("<%dQ" % (ns * self.ncls), raw, h["off_trans"]) if ns else ()
        self.sinfo = struct.unpack_from("<%dQ" % ns, raw, h["off_sinfo"]) if ns else ()
        self.pool = raw[h["off_pool"]:h["off_pool"] + h["pool_bytes"]]
        self.nft = [struct.unpack_from("<BBBBii", raw, h["off_nft"] + 12 * k) for k in range(h["nft_states"])]
        words = struct.unpack_from("<70I", raw, 0)
        for k, name in EXT.items():
            h[name] = words[k]

    # ---- table-free byte map (k_bytemap_swar): the LUT rebuilt from the ranges the kernel is given ----------
    def swar_lut(self):
        h = self.h
        if not h["swar_n"]:
            return None
        lut = bytearray(range(256))
        for k in range(h["swar_n"]):
            w = h["swar%d" % k]
            lo, hi, delta = w & 0xFF, (w >> 8) & 0xFF, (w >> 16) & 0xFF
            assert lo <= hi and (lo < 0x80) == (hi < 0x80)
            for c in range(256):
                # the kernel's test: 7-bit compare inside the selected half of the byte space
                if (c < 0x80) == (lo < 0x80) and (lo & 0x7F) <= (c & 0x7F) <= (hi & 0x7F):
                    lut[c] = (c + delta) & 0xFF
        return bytes(lut)

    # ---- gram tables (k_gram): position-parallel form of the scan over a REGULAR buffer ----------------------
    def scan_vmap(self, data: bytes) -> bytes:
        """What k_bmv_* compute (bmv_kernel.cuh): every byte prints the string of its vent entry - except, in the RUN form,
        a byte whose predecessor has the same non-zero run class (bits 28..31): it is swallowed by the match that started
        the run.  Regular buffers only."""
        h = self.h
        assert h["off_vent"] and data.endswith(b"\n") and b"\0" not in data
        vent = struct.unpack_from("<256I", self.raw, h["off_vent"])
        run = h["vrun_classes"] > 0
        out = bytearray()
        prev_rc = 0
        for c in data:
            e = vent[c]
            rc = (e >> 28) if run else 0
            if not (rc and rc == prev_rc):
                ln = e & 0xFF
                if ln == 1:
                    out.append((e >> 8) & 0xFF)
                elif ln > 1:
                    off = h["off_vpool"] + (((e >> 16) & 0xFFF) if run else (e >> 16))
                    out += self.raw[off:off + ln]
            prev_rc = rc
        return bytes(out)

    def scan_litv(self, data: bytes) -> bytes:
        """What k_litv_count / k_bmv_write<BMV_LIT> compute: wherever a word stands its lit_len bytes are replaced by the
        word's string (gpool, any length); every other byte prints itself.  Regular buffers only."""
        h = self.h
        nw, L = h["lit_words"], h["lit_len"]
        assert nw and data.endswith(b"\n") and b"\0" not in data
        out = bytearray()
        p = 0
        while p < len(data):
            hit = None
            if p + L <= len(data):
                for w in range(nw):
                    slots = [h["lit%d" % (w * L + k)] for k in range(L)]
                    if all((s & 0xFF) <= data[p + k] <= ((s >> 8) & 0xFF) for k, s in enumerate(slots)):
                        hit = w
            if hit is None:
                out.append(data[p])
                p += 1
            else:
                o = h["lit_out%d" % hit]
                off = h["off_gpool"] + (o & 0xFFFFF)
                out += self.raw[off:off + ((o >> 20) & 0xFF)]
                p += L
        return bytes(out)

    def scan_lit(self, data: bytes) -> bytes:
        """What k_lit computes (lit_kernel.cuh): wherever one of the lit_words words stands - position k of a word accepts
        one byte range - its lit_len bytes are replaced by what it prints.  No line structure, no "inside an earlier
        match" bookkeeping: the host only emits this form when words cannot overlap.  Regular buffers only."""
        h = self.h
        nw, L = h["lit_words"], h["lit_len"]
        assert nw and data.endswith(b"\n") and b"\0" not in data
        out = bytearray(data)
        for p in range(len(data) - L + 1):
            for w in range(nw):
                slots = [h["lit%d" % (w * L + k)] for k in range(L)]
                if all((s & 0xFF) <= data[p + k] <= ((s >> 8) & 0xFF) for k, s in enumerate(slots)):
                    for k, s in enumerate(slots):
                        out[p + k] = (s >> 16) & 0xFF
        return bytes(out)

    def scan_gram(self, data: bytes) -> bytes:
        """What k_gram computes: a match starts at every position whose D-gram bit is set and that no earlier
        match of the line covers; it prints gout.  Regular buffers only (no NUL, final newline)."""
        h = self.h
        assert h["flags"] & PF_GRAM and data.endswith(b"\n") and b"\0" not in data
        D, ncls, n = h["gram_depth"], self.ncls, len(data)
        bits = struct.unpack_from("<%dI" % ((h["gram_entries"] + 31) // 32), self.raw, h["off_gbits"])
        gout = struct.unpack_from("<%dI" % h["gram_entries"], self.raw, h["off_gout"])
        gpool = self.raw[h["off_gpool"]:h["off_gpool"] + h["gpool_bytes"]]
        cm = self.classmap
        out = bytearray(data)
        covered_until = 0
        for p in range(n):
            if p < covered_until:
                continue
            idx = 0
            for i in range(D - 1, -1, -1):
                c = data[p + i] if p + i < n else 10
                idx = idx * ncls + cm[c]
            if not (bits[idx >> 5] >> (idx & 31)) & 1:
                continue
            g = gout[idx]
            if h["gram_flags"] & 1:
                m = h["gram_len"]
                rep = bytes((g >> (8 * k)) & 0xFF for k in range(m))
            else:
                m = (g >> 28) + 1
                ln = (g >> 20) & 0xFF
                rep = gpool[g & 0xFFFFF:(g & 0xFFFFF) + ln]
                assert h["gram_len"] in (0, m)
            assert len(rep) == m, "k_gram serves length-preserving programs"
            out[p:p + m] = rep
            covered_until = p + m
        return bytes(out)

    def ref(self, hi):
        ln, off = hi >> 24, hi & 0xFFFFFF
        if ln == 255:
            ln = struct.unpack_from("<I", self.pool, off)[0]
            off += 4
        return self.pool[off:off + ln]

    # one attempt over payload[p:]; -> (consumed or -1, output bytes)
    def attempt(self, payload, p):
        h = self.h
        s, i, acc = 0, 0, b""
        m = 0 if h["start_info"] & S_HAS_FINAL else -1
        decided = bool(h["start_info"] & S_DECIDED)
        while not decided and p + i < len(payload):
            e = self.trans[s * self.ncls + self.classmap[payload[p + i]]]
            lo, hi = e & 0xFFFFFFFF, e >> 32
            if lo & 0xFFFFFF == DEAD:
                break
            acc += self.ref(hi)
            s = lo & 0xFFFFFF
            i += 1
            if lo & E_MARK:
                m = i
            if lo & E_DECIDED:
                decided = True
        si = self.sinfo[s]
        sf = si & 0xFFFFFFFF
        if sf & S_DIVERGE:
            raise Diverged()
        if not sf & S_HAS_FINAL:
            return -1, b""
        return m, acc + self.ref(si >> 32)

    def attempt_backtrack(self, payload, p):
        nft, h = self.nft, self.h
        s, i, o, stack = h["nft_start"], 0, b"", []
        n = len(payload) - p
        steps = 0
        while stack or s >= 0:
            steps += 1
            if steps > 2_000_000 or len(stack) > 65536 or len(o) > 1 << 16:
                raise Diverged()
            if s < 0:
                s, i, o = stack.pop()
                if s < 0:
                    continue
            op, val, _, _, a, b = nft[s]
            if op == 1:
                if i < n and payload[p + i] == val:
                    i += 1
                    s = a
                else:
                    s = -1
            elif op == 0:
                o += bytes([val]); s = a
            elif op == 2:
                stack.append((a, i, o)); s = b
            elif op == 3:
                stack.append((b, i, o)); s = a
            elif op == 4:
                s = a
            else:
                z = o.find(b"\0")
                return i, (o if z < 0 else o[:z])
        return -1, b""

    # TRRE_EXEC_LAZY: rows by raw byte from the on-demand builder (Program.lazy_view / lazy_build), as the lazy
    # kernel walks them; a missing row is built on the spot (the GPU reports it and repeats the pass instead)
    def attach_lazy(self, prog):
        self.lazy = prog
        self.lazy_builds = 0

    def attempt_lazy(self, payload, p):
        import ctypes
        prog = self.lazy
        while True:
            n, rows_a, sinfo_a, pool_a, pool_n = prog.lazy_view()
            rows = (ctypes.c_uint64 * (n * 256)).from_address(rows_a)
            sinfo = (ctypes.c_uint64 * n).from_address(sinfo_a)
            pool = ctypes.string_at(pool_a, pool_n) if pool_n else b""

            def ref(hi):
                ln, off = hi >> 24, hi & 0xFFFFFF
                if ln == 255:
                    ln = struct.unpack_from("<I", pool, off)[0]
                    off += 4
                return pool[off:off + ln]

            s, i, acc, m, miss = 0, 0, b"", -1, None
            if sinfo[0] & S_HAS_FINAL:
                m = 0
            decided = bool(sinfo[0] & S_DECIDED)
            while not decided and p + i < len(payload):
                e = rows[s * 256 + payload[p + i]]
                lo, hi = e & 0xFFFFFFFF, e >> 32
                if lo & 0xFFFFFF == 0xFFFFFE:
                    miss = s
                    break
                if lo & 0xFFFFFF == 0xFFFFFD:
                    raise Diverged()
                if lo & 0xFFFFFF == DEAD:
                    break
                acc += ref(hi)
                s = lo & 0xFFFFFF
                i += 1
                if lo & E_MARK:
                    m = i
                if lo & E_DECIDED:
                    decided = True
            if miss is not None:
                prog.lazy_build([miss], 0)
                self.lazy_builds += 1
                continue
            sf = sinfo[s] & 0xFFFFFFFF
            if sf & S_DIVERGE:
                raise Diverged()
            if not sf & S_HAS_FINAL:
                return -1, b""
            return m, acc + ref(sinfo[s] >> 32)

    def scan(self, data: bytes) -> bytes:
        """The reference's scan driver over getline records (trre_nft.c:775-790), tables from the blob."""
        att = {2: self.attempt_backtrack, 3: self.attempt_lazy}.get(self.h["exec_kind"], self.attempt)
        out = bytearray()
        pos, n = 0, len(data)
        while pos < n:
            nl = data.find(b"\n", pos)
            rec = data[pos:(nl + 1 if nl >= 0 else n)]
            pos += len(rec)
            payload = rec[:-1]
            z = payload.find(b"\0")
            if z >= 0:
                payload = payload[:z]
            p = 0
            while True:
                k, o = att(payload, p)
                out += o
                if p >= len(payload):
                    break
                if k > 0:
                    p += k
                else:
                    out.append(payload[p]); p += 1
            out.append(10)
        return bytes(out)
