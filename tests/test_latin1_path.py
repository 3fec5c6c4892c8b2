"""Latin-1 rows of mixed-script columns (csrc/row_ascii_reg.cuh: transcode_latin1 + 8 bit planes), host build.

Rows whose characters are all below U+0100 are transcoded from UTF-8 to one byte per character (in place)
and then run the bit-plane path with 8 planes.  Results must match the oracle on the ORIGINAL strings.
"""
import ctypes
import random

from test_pair_algos import algos, check_multi  # noqa: F401  (algos is a fixture)


def latin_chars():
    return [chr(c) for c in range(0xC0, 0x100) if c not in (0xD7, 0xF7)] + [chr(0x80), chr(0xA9), chr(0xBF)]


def test_latin1_transcoded_rows(algos, oracle):  # noqa: F811
    from oracle.oracle import _pack

    algos.algos_batch_latin1_multi.restype = ctypes.c_int
    algos.algos_batch_latin1_multi.argtypes = [ctypes.c_int, ctypes.c_int64] + [ctypes.c_void_p] * 6
    rng = random.Random(2024)
    latin = latin_chars()
    ascii_part = "abcdefghijklmnopqrstuvwxyz-' "
    e_acute, y_diaeresis, a_grave, thorn = chr(0xE9), chr(0xFF), chr(0xC0), chr(0xFE)

    def word(n):
        return "".join(rng.choice(latin) if rng.random() < 0.3 else rng.choice(ascii_part) for _ in range(n))

    a, b = [], []
    while len(a) < 12000:
        x = word(rng.randint(0, 24))
        if rng.random() < 0.7:
            y = list(x)
            for _ in range(rng.randint(0, 3)):
                op, pos = rng.randint(0, 3), rng.randint(0, len(y))
                if op == 0 and y:
                    y[min(pos, len(y) - 1)] = word(1)
                elif op == 1:
                    y.insert(pos, word(1))
                elif op == 2 and y:
                    del y[min(pos, len(y) - 1)]
                elif op == 3 and len(y) > 1:
                    q = min(pos, len(y) - 2)
                    y[q], y[q + 1] = y[q + 1], y[q]
            y = "".join(y)
        else:
            y = word(rng.randint(0, 24))
        if len(x.encode()) <= 32 and len(y.encode()) <= 32:
            a.append(x)
            b.append(y)
    for la in (0, 1, 2, 15, 16):  # 16 two-byte characters fill the 32 bytes exactly
        for lb in (0, 1, 2, 15, 16):
            a += [e_acute * la, e_acute * la, "a" * la]
            b += [e_acute * lb, y_diaeresis * lb, e_acute * lb]
    a += [("a" + e_acute) * 10 + "b", a_grave + y_diaeresis, "x" * 32, "abc" + e_acute]
    b += [(e_acute + "a") * 10 + "b", a_grave + thorn, "x" * 30 + e_acute, "abc" + e_acute]
    ad, ao, _ = _pack(a)
    bd, bo, _ = _pack(b)
    check_multi(oracle, lambda g, ints, vals: algos.algos_batch_latin1_multi(
        g, len(a), ad.ctypes.data, ao.ctypes.data, bd.ctypes.data, bo.ctypes.data, ints.ctypes.data,
        vals.ctypes.data), a, b)
