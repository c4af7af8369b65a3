"""Writes the Julia binding shown in INTEGRATION.md out as source files (julia/algorithms_b200.jl, julia/ext/IntegralsB200Ext.jl), so that
the document and the deliverable cannot drift apart (tests/test_cabi.py re-runs this and compares).  Julia is not installed in this
image: the files are delivered as source, untested."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = ("# GENERATED from INTEGRATION.md (section {sec}) by scripts/extract_julia_binding.py -- edit the document, not this file.\n"
          "# Binding of libintegrals_b200.so (C ABI: include/integrals_b200.h) into Integrals.jl; untested here (no Julia in the build image).\n\n")


def blocks():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    out = {}
    for sec, title in (("2", "## 2."), ("3", "## 3.")):
        start = text.index(title)
        m = re.search(r"```julia\n(.*?)```", text[start:], flags=re.S)
        out[sec] = m.group(1)
    return out


def render():
    b = blocks()
    return {os.path.join("julia", "algorithms_b200.jl"): HEADER.format(sec="2: algorithm structs and registered integrands, next to src/algorithms_extension.jl") + b["2"],
            os.path.join("julia", "ext", "IntegralsB200Ext.jl"): HEADER.format(sec="3: the package extension") + b["3"]}


if __name__ == "__main__":
    for rel, src in render().items():
        path = os.path.join(ROOT, rel)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        open(path, "w").write(src)
        print("wrote", rel, len(src.splitlines()), "lines")
