#!/usr/bin/env python
"""spv_constants.py — lists the scalar constants of a SPIR-V module (TEST INFRASTRUCTURE ONLY).

The reference executes shaders/*.spv, not the GLSL text (main.cpp:541-543). The pixel oracle is pinned against the
TEXT (oracle/_ref/libref_shade.so); this decoder closes the remaining gap by checking that the shipped binaries carry
the constants the text implies (spp 32, depth bound 8, tmin, tmax, camera, sky, 2*pi, 1/(2*pi), pi, 2^-32, the PCG
multipliers ...). It reads the .spv files where they lie and prints/returns values only.

    python oracle/spv_constants.py /root/reference/shaders/raygen.rgen.spv
"""
import struct
import sys

OP_TYPE_INT, OP_TYPE_FLOAT, OP_CONSTANT = 21, 22, 43


def constants(path):
    """{'f32': sorted bit patterns of the float constants, 'u32': sorted values of the 32-bit integer constants}"""
    data = open(path, "rb").read()
    words = struct.unpack("<%dI" % (len(data) // 4), data)
    assert words[0] == 0x07230203, "not a SPIR-V module"
    types, out, i = {}, {"f32": set(), "u32": set()}, 5
    while i < len(words):
        wc, op = words[i] >> 16, words[i] & 0xFFFF
        if op == OP_TYPE_INT and words[i + 2] == 32:
            types[words[i + 1]] = "u32"
        elif op == OP_TYPE_FLOAT and words[i + 2] == 32:
            types[words[i + 1]] = "f32"
        elif op == OP_CONSTANT and words[i + 1] in types:
            out[types[words[i + 1]]].add(words[i + 3])
        i += wc
    return {k: sorted(v) for k, v in out.items()}


if __name__ == "__main__":
    for p in sys.argv[1:]:
        c = constants(p)
        print(p)
        print("  f32:", ", ".join("%s(0x%08x)" % (struct.unpack("<f", struct.pack("<I", b))[0], b) for b in c["f32"]))
        print("  u32:", ", ".join(str(v) for v in c["u32"]))
