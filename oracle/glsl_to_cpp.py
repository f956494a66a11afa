#!/usr/bin/env python
"""glsl_to_cpp.py — mechanical token transform that lets a C++ compiler read the reference's GLSL shader text.

TEST INFRASTRUCTURE ONLY. Reads shaders/{common.glsl,raygen.rgen,closesthit.rchit,miss.rmiss} from the reference
checkout WHERE THEY LIE and writes `<name>.inc` files into oracle/_ref/gen/ (git-ignored, never committed): the
reference's sources are not copied into this repository. oracle/ref_shade_glue.cpp includes the results under
oracle/glsl_shim.h. No statement, expression, constant or operation order of the shader text is changed; the rules
below only re-spell what C++ cannot parse:

  R1  `#version` / `#extension` lines                   -> commented out
  R2  `#include "x"`                                    -> `#include "x.inc"`
  R3  parameter qualifiers: `inout T n` / `out T n`     -> `T& n`;  `in T n` -> `T n`
  R4  float literals without suffix (`0.001`, `2.0`)    -> `0.001f`, `2.0f` (a GLSL literal is a 32-bit float; a C++
                                                           one would be a double and change the arithmetic)
  R5  anonymous interface blocks:
        `uniform Name { members };`                      -> `members` at namespace scope (GLSL block members without an
                                                           instance name are global names)
        `buffer Name { T name[]; };`                     -> `glsl::buffer_array<T> name;`
  R6  every call / constructor with >= 2 arguments      -> GLSL_CALL(f, args...) (left-to-right evaluation, see the shim)
  R7  the two loop bounds stay overridable at run time: `int maxSamples = 32;` -> `int maxSamples = glsl::ref_spp(32);`
      and `depth < 8` -> `depth < glsl::ref_depth(8)`; without an override they return the literal of the text

    python oracle/glsl_to_cpp.py /root/reference/shaders oracle/_ref/gen
"""
import os
import re
import sys

FILES = ["common.glsl", "raygen.rgen", "closesthit.rchit", "miss.rmiss"]
KEYWORDS = {"if", "for", "while", "switch", "return", "layout", "sizeof", "else", "do"}
IDENT = re.compile(r"[A-Za-z_]\w*")


def match_paren(s, i):
    """index of the ')' matching the '(' at s[i] (comment-aware)."""
    depth = 0
    while i < len(s):
        if s.startswith("//", i):
            i = s.index("\n", i)
            continue
        c = s[i]
        if c == "(":
            depth += 1
        elif c == ")":
            depth -= 1
            if depth == 0:
                return i
        i += 1
    raise ValueError("unbalanced parentheses")


def split_args(s):
    """top-level comma split of an argument list (comments stay inside the argument they follow or precede)."""
    args, depth, start, i = [], 0, 0, 0
    while i < len(s):
        if s.startswith("//", i):
            i = s.index("\n", i) if "\n" in s[i:] else len(s)
            continue
        c = s[i]
        if c in "([{":
            depth += 1
        elif c in ")]}":
            depth -= 1
        elif c == "," and depth == 0:
            args.append(s[start:i])
            start = i + 1
        i += 1
    tail = s[start:]
    if tail.strip() or args:
        args.append(tail)
    return args


def wrap_calls(s):
    """R6 over a chunk of source text."""
    out, i, prev = [], 0, ""
    while i < len(s):
        if s.startswith("//", i):
            j = s.index("\n", i) if "\n" in s[i:] else len(s)
            out.append(s[i:j]); i = j
            continue
        if s[i] == "#":  # preprocessor line: verbatim
            j = s.index("\n", i) if "\n" in s[i:] else len(s)
            out.append(s[i:j]); i = j
            continue
        m = IDENT.match(s, i)
        if m:
            name, j = m.group(0), m.end()
            k = j
            while k < len(s) and s[k] in " \t":
                k += 1
            if k < len(s) and s[k] == "(" and name not in KEYWORDS:
                close = match_paren(s, k)
                inner = s[k + 1:close]
                if prev and IDENT.fullmatch(prev) and prev not in KEYWORDS:
                    out.append(name + s[j:k] + "(" + inner + ")")  # a declaration `T name(params)`: verbatim
                else:
                    args = [wrap_calls(a) for a in split_args(inner)]
                    if len(args) >= 2:
                        out.append("GLSL_CALL(" + name + "," + ",".join(args) + ")")
                    else:
                        out.append(name + s[j:k] + "(" + ",".join(args) + ")")
                i, prev = close + 1, ")"
                continue
            out.append(name); i, prev = j, name
            continue
        if not s[i].isspace():
            prev = s[i]
        out.append(s[i]); i += 1
    return "".join(out)


def transform(text):
    text = re.sub(r"^(#version.*|#extension.*)$", r"// \1", text, flags=re.M)                      # R1
    text = re.sub(r'^#include "([^"]+)"', r'#include "\1.inc"', text, flags=re.M)                  # R2
    text = re.sub(r"\binout\s+(\w+)\s+(\w+)", r"\1& \2", text)                                      # R3
    text = re.sub(r"\bout\s+(\w+)\s+(\w+)", r"\1& \2", text)
    text = re.sub(r"\bin\s+(\w+)\s+(\w+)", r"\1 \2", text)
    text = re.sub(r"(?<![\w.])(\d+\.\d*|\.\d+)([eE][+-]?\d+)?(?![\w.])", r"\1\2f", text)            # R4
    text = re.sub(r"\bbuffer\s+\w+\s*\{\s*(\w+)\s+(\w+)\s*\[\s*\]\s*;\s*\}\s*;", r"glsl::buffer_array<\1> \2;", text)  # R5
    text = re.sub(r"\buniform\s+\w+\s*\{([^}]*)\}\s*;", lambda m: m.group(1).strip(), text)
    text = wrap_calls(text)                                                                        # R6
    text, n1 = re.subn(r"\bint maxSamples = (\d+);", r"int maxSamples = glsl::ref_spp(\1);", text)  # R7
    text, n2 = re.subn(r"\bdepth < (\d+)\b", r"depth < glsl::ref_depth(\1)", text)
    return text, n1, n2


def main(src_dir, dst_dir):
    os.makedirs(dst_dir, exist_ok=True)
    for name in FILES:
        with open(os.path.join(src_dir, name)) as f:
            text = f.read()
        out, n1, n2 = transform(text)
        if name == "raygen.rgen" and (n1, n2) != (1, 1):
            raise SystemExit(f"{name}: expected one `int maxSamples = N;` and one `depth < N` (found {n1}, {n2})")
        with open(os.path.join(dst_dir, name + ".inc"), "w") as f:
            f.write(f"// GENERATED by oracle/glsl_to_cpp.py from {os.path.join(src_dir, name)} -- do not commit\n" + out)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
