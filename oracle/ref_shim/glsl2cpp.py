#!/usr/bin/env python3
"""TEST INFRASTRUCTURE (oracle/): turns one of the reference's compute shaders into a C++ include file, at build time, from the source where
it lies under /root/reference (nothing of the shader is committed to this repository; the output goes to the git-ignored oracle/_ref/glsl/).

    glsl2cpp.py <MeshFragments root> <shader path relative to it> <out.inc>

What is rewritten — declarations only, the statements of the shader are compiled as they are:
  * `#include <file>` is replaced by the file's text, recursively, as ShaderProgram::includeLibraries does (ShaderProgram.cpp:362-404);
  * `#version`, `#extension` and the `layout (local_size_variable) in;` line are dropped;
  * `layout (std430, binding = N) buffer B { T name[]; };`  ->  `glsl_buffer<T> name;` (the driver binds a host array to it)
    `layout (std430, binding = N) buffer B { T name; };`    ->  `T* p_name;` + `#define name (*p_name)` (undefined again at the end)
  * `subroutine R type(args);` -> a function-pointer typedef, `subroutine uniform type name;` -> a variable of it, `subroutine(type)` dropped;
  * `void main()` -> `void shader_main()`;  `.xyz` -> `.xyz()` (C++ has no swizzles);
  * `T name = ... name(...)` (a variable named like the function its initialiser calls: legal GLSL, the variable is not in scope yet;
    in C++ it is) -> the variable and its later uses in that block become `name_v` (marchingCubes-comp.glsl:147-150).
`uniform` and the parameter qualifier `in` are emptied by macros in the including file (oracle/ref_shim/ref_glsl.cpp)."""
import os
import re
import sys


def expand(root, rel, seen=()):
    text = open(os.path.join(root, rel)).read()
    out = []
    for line in text.splitlines():
        m = re.match(r"\s*#include\s*<([^>]+)>", line)
        if m:
            out.append(expand(root, m.group(1), seen + (rel,)))
        else:
            out.append(line)
    return "\n".join(out)


def rename_shadowing(src):
    lines = src.splitlines()
    i = 0
    while i < len(lines):
        m = re.match(r"(\s*)(\w+)\s+(\w+)\s*=\s*(.*)$", lines[i])
        if m and re.search(r"\b" + re.escape(m.group(3)) + r"\s*\(", m.group(4)):
            name, depth = m.group(3), 0
            lines[i] = f"{m.group(1)}{m.group(2)} {name}_v = {m.group(4)}"
            for j in range(i + 1, len(lines)):
                depth += lines[j].count("{") - lines[j].count("}")
                if depth < 0:
                    break
                lines[j] = re.sub(r"\b" + re.escape(name) + r"\b(?!\s*\()", name + "_v", lines[j])
        i += 1
    return "\n".join(lines)


def translate(src):
    out, undef = [], []
    src = rename_shadowing(src)
    for line in src.splitlines():
        s = line.strip()
        if s.startswith("#version") or s.startswith("#extension"):
            continue
        if re.match(r"layout\s*\(\s*local_size_variable\s*\)\s*in\s*;", s):
            continue
        m = re.match(r"layout\s*\([^)]*\)\s*buffer\s+\w+\s*\{\s*(\w+)\s+(\w+)\s*(\[\s*\])?\s*;\s*\}\s*;", s)
        if m:
            ty, name, arr = m.groups()
            if arr:
                out.append(f"glsl_buffer<{ty}> {name};")
            else:
                out.append(f"{ty}* p_{name};")
                out.append(f"#define {name} (*p_{name})")
                undef.append(name)
            continue
        m = re.match(r"subroutine\s+uniform\s+(\w+)\s+(\w+)\s*;", s)
        if m:
            out.append(f"{m.group(1)} {m.group(2)};")
            continue
        m = re.match(r"subroutine\s+(\w+)\s+(\w+)\s*\((.*)\)\s*;", s)
        if m:
            out.append(f"typedef {m.group(1)} (*{m.group(2)})({m.group(3)});")
            continue
        if re.match(r"subroutine\s*\(\s*\w+\s*\)\s*$", s):
            continue
        line = re.sub(r"\bvoid\s+main\s*\(\s*\)", "void shader_main()", line)
        line = re.sub(r"\.xyz\b", ".xyz()", line)
        out.append(line)
    out += [f"#undef {n}" for n in undef]
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    root, rel, dst = sys.argv[1:4]
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    with open(dst, "w") as f:
        f.write(f"// generated at build time from {rel} (reference source, not committed)\n")
        f.write(translate(expand(root, rel)))
