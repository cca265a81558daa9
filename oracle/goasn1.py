"""ORACLE (test infrastructure, NOT product code).

Restatement of the subset of Go 1.14 ``encoding/asn1`` that the reference uses at
``bccsp/utils/ecdsa.go:43-67``: ``asn1.Unmarshal(raw, &ECDSASignature{R,S *big.Int})``
and ``asn1.Marshal`` of the same struct (``bccsp/utils/ecdsa.go:39-41``).

Rules restated ("Go encoding/asn1 rules"):
  * DER only: definite, minimal lengths; long form < 0x80 rejected; leading zero
    length bytes rejected; > 4-byte lengths / >= 2^23 rejected;
  * outer element must be universal, constructed, tag 16 (SEQUENCE);
  * each field must be universal, primitive, tag 2 (INTEGER), non-empty and
    minimally encoded; negative values parse (two's complement);
  * bytes after the second INTEGER inside the SEQUENCE, and bytes after the
    SEQUENCE, are accepted and ignored.
"""


class Asn1Error(Exception):
    pass


def _parse_base128(b, off):
    ret = 0
    shifted = 0
    while off < len(b):
        if shifted == 5:
            raise Asn1Error("base 128 integer too large")
        ret <<= 7
        c = b[off]
        if shifted == 0 and c == 0x80:
            raise Asn1Error("integer is not minimally encoded")
        ret |= c & 0x7F
        off += 1
        shifted += 1
        if c & 0x80 == 0:
            if ret > (1 << 31) - 1:
                raise Asn1Error("base 128 integer too large")
            return ret, off
    raise Asn1Error("truncated base 128 integer")


def _tag_and_length(b, off):
    """Go parseTagAndLength."""
    if off >= len(b):
        raise Asn1Error("parseTagAndLength should not be called without at least a single byte to read")
    c = b[off]
    off += 1
    cls = c >> 6
    compound = bool(c & 0x20)
    tag = c & 0x1F
    if tag == 0x1F:
        tag, off = _parse_base128(b, off)
        if tag < 0x1F:
            raise Asn1Error("non-minimal tag")
    if off >= len(b):
        raise Asn1Error("truncated tag or length")
    c = b[off]
    off += 1
    if c & 0x80 == 0:
        length = c & 0x7F
    else:
        nbytes = c & 0x7F
        if nbytes == 0:
            raise Asn1Error("indefinite length found (not DER)")
        length = 0
        for _ in range(nbytes):
            if off >= len(b):
                raise Asn1Error("truncated tag or length")
            c = b[off]
            off += 1
            if length >= 1 << 23:
                raise Asn1Error("length too large")
            length = (length << 8) | c
            if length == 0:
                raise Asn1Error("superfluous leading zeros in length")
        if length < 0x80:
            raise Asn1Error("non-minimal length")
    return cls, compound, tag, length, off


def _parse_bigint_field(b, off):
    if off == len(b):
        raise Asn1Error("sequence truncated")
    cls, compound, tag, length, off = _tag_and_length(b, off)
    if off + length > len(b):
        raise Asn1Error("data truncated")
    if cls != 0 or tag != 2 or compound:
        raise Asn1Error("tags don't match")
    body = b[off:off + length]
    if len(body) == 0:
        raise Asn1Error("empty integer")
    if len(body) > 1 and ((body[0] == 0 and body[1] & 0x80 == 0) or (body[0] == 0xFF and body[1] & 0x80 == 0x80)):
        raise Asn1Error("integer not minimally-encoded")
    return int.from_bytes(body, "big", signed=True), off + length


def unmarshal_ecdsa_signature(raw: bytes):
    """asn1.Unmarshal(raw, &ECDSASignature{}) -> (R, S) as Python ints (may be <= 0).
    Raises Asn1Error where Go returns an error."""
    b = bytes(raw) if raw is not None else b""
    if len(b) == 0:
        raise Asn1Error("sequence truncated")
    cls, compound, tag, length, off = _tag_and_length(b, 0)
    if off + length > len(b):
        raise Asn1Error("data truncated")
    if cls != 0 or tag != 16 or not compound:
        raise Asn1Error("tags don't match")
    inner = b[off:off + length]
    r, ioff = _parse_bigint_field(inner, 0)
    s, ioff = _parse_bigint_field(inner, ioff)
    return r, s


def _der_int(v: int) -> bytes:
    n = max(1, (v.bit_length() + 8) // 8) if v >= 0 else max(1, ((-v - 1).bit_length() + 8) // 8)
    return v.to_bytes(n, "big", signed=True)


def _der_len(n: int) -> bytes:
    if n < 0x80:
        return bytes([n])
    body = n.to_bytes((n.bit_length() + 7) // 8, "big")
    return bytes([0x80 | len(body)]) + body


def marshal_ecdsa_signature(r: int, s: int) -> bytes:
    """asn1.Marshal(ECDSASignature{r, s}) (bccsp/utils/ecdsa.go:39-41)."""
    ri, si = _der_int(r), _der_int(s)
    body = b"\x02" + _der_len(len(ri)) + ri + b"\x02" + _der_len(len(si)) + si
    return b"\x30" + _der_len(len(body)) + body
