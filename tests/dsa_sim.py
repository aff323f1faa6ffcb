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

    def scan(self, data: bytes) -> bytes:
        """The reference's scan driver over getline records (trre_nft.c:775-790), tables from the blob."""
        att = self.attempt_backtrack if self.h["exec_kind"] == 2 else self.attempt
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
