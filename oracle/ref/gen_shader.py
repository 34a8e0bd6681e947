#!/usr/bin/env python
"""gen_shader.py — makes the reference's fragment shader #include-able by g++.

    python gen_shader.py /root/reference/fragment.shd oracle/_ref/fragment.gen.inc

TEST INFRASTRUCTURE (oracle/_ref recipe).  Reads the shader FROM THE REFERENCE TREE and
writes a derived file into oracle/_ref/ (git-ignored; no reference source enters the
repository).  The only edit is the one C++ cannot express with a macro: GLSL's `out`
parameter qualifier becomes a C++ reference,

        , out float tmin      ->      , float& tmin

Everything else compiles as written, given oracle/ref/glsl_compat.hpp and these macros set
by ref_driver.cpp:  `uniform` -> static, the fragment output `out vec4 frag_color;` ->
static thread_local, `main` -> shader_main, GLSL float literals -> -fsingle-precision-constant.
The per-variant `#define` lines that ShaderRendering.hs:119-123 prepends are passed as -D.
"""
import re
import sys
from pathlib import Path

# an `out` qualifier inside a parameter list follows '(' or ','
OUT_PARAM = re.compile(r"([,(]\s*)out\s+(float|vec2|vec3|vec4)\s+")


def convert(text: str):
    new, n = OUT_PARAM.subn(lambda m: "%s%s& " % (m.group(1), m.group(2)), text)
    return new, n


def main(src: str, dst: str) -> int:
    text = Path(src).read_text(encoding="utf-8")
    new, n = convert(text)
    # the shader has 14 `out` parameters (fragment.shd:42,258-260,323,599-600,620-621,844-845)
    # and one `out` fragment output (:18), which stays for the macro
    left = len(re.findall(r"\bout\b", re.sub(r"//[^\n]*", "", new)))
    if n != 14 or left != 1:
        sys.stderr.write("gen_shader: unexpected shader text (%d out parameters rewritten, %d left)\n" % (n, left))
        return 1
    assert new.count("\n") == text.count("\n")          # line numbers stay those of fragment.shd
    Path(dst).parent.mkdir(parents=True, exist_ok=True)
    Path(dst).write_text("#line 1 \"fragment.shd\"\n" + new, encoding="utf-8")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1], sys.argv[2]))
