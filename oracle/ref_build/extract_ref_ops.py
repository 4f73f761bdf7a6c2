#!/usr/bin/env python3
"""Build step of oracle/_ref (TEST INFRASTRUCTURE): pull the reference's OWN method bodies out of its
source files, verbatim, into generated include files under oracle/_ref/gen/ (git-ignored build output).

The reference's kaldi-matrix.cc / cu-matrix.cc / nnet-loss.cc cannot be compiled as whole translation
units here: they include a dozen upstream Kaldi headers that are not vendored.  The methods the
LstmProjectedStreams path calls are self-contained, though, so this script copies exactly those
definitions (template line(s) + signature + body, located by qualified name and brace matching) from the
files where they lie under /root/reference into .inc files, and oracle/ref_build/shim/kaldi-ref-shim.h
includes them inside its own minimal MatrixBase / CuMatrixBase class declarations.  Nothing from the
reference is committed to this repository; the generated files exist only on a machine that has
/root/reference mounted.

usage: extract_ref_ops.py <reference-root> <out-dir>
"""
import os
import re
import sys

# (source file, qualified method name, occurrence index) -> goes to <out>/<group>.inc
WANTED = {
    "km_ops": ("google/matrix/kaldi-matrix.cc", [
        "MatrixBase<Real>::AddMatMat",      # :159-175   cblas_Xgemm contract
        "MatrixBase<Real>::AddMat",         # :345       axpy per row
        "MatrixBase<Real>::AddMatDiagVec",  # :447-473   <jiayu>
        "MatrixBase<Real>::AddMatDotMat",   # :475-497   <jiayu>
        "MatrixBase<Real>::Scale",          # :1033
        "MatrixBase<Real>::SetZero",        # :1123
        "MatrixBase<Real>::Add",            # :1452
        "MatrixBase<Real>::ApplyFloor",     # :1868
        "MatrixBase<Real>::ApplyCeiling",   # :1878
        "MatrixBase<Real>::Tanh",           # :2457
        "MatrixBase<Real>::Sigmoid",        # :2545
        "MatrixBase<Real>::DiffSigmoid",    # :2561
        "MatrixBase<Real>::DiffTanh",       # :2578
        "MatrixBase<Real>::AddVecToRows",   # :2596
        "MatrixBase<Real>::MulElements",    # :977    (Xent)
        "MatrixBase<Real>::Sum",            # :1007   (Xent)
        "MatrixBase<Real>::MulRowsVec",     # :1048   (Xent)
        "MatrixBase<Real>::ApplyLog",       # :1889   (Xent)
    ]),
    "cum_ops": ("google/cudamatrix/cu-matrix.cc", [
        "CuMatrixBase<Real>::AddMat",         # :796-821
        "CuMatrixBase<Real>::AddVecToRows",   # :878-902
        "CuMatrixBase<Real>::AddMatMat",      # :909-945
        "CuMatrixBase<Real>::AddMatDiagVec",  # :1014-1046  <jiayu>
        "CuMatrixBase<Real>::AddMatDotMat",   # :1048-1068  <jiayu>
        "CuMatrixBase<Real>::Sigmoid",        # :1072-1091
        "CuMatrixBase<Real>::DiffSigmoid",    # :1221-1241
        "CuMatrixBase<Real>::Tanh",           # :1244-1263
        "CuMatrixBase<Real>::DiffTanh",       # :1267-1286
        "CuMatrixBase<Real>::ApplyFloor",     # :1752-1768
        "CuMatrixBase<Real>::ApplyCeiling",   # :1770-1786
        "CuMatrixBase<Real>::Add",            # :511    (Xent)
        "CuMatrixBase<Real>::ApplyLog",       # :590    (Xent)
        "CuMatrixBase<Real>::MulElements",    # :610    (Xent)
        "CuMatrixBase<Real>::MulRowsVec",     # :682    (Xent)
        "CuMatrixBase<Real>::FindRowMaxId",   # :1289   (Xent)
        "CuMatrixBase<Real>::Sum",            # :1984   (Xent)
    ]),
    "loss_ops": ("google/nnet/nnet-loss.cc", [
        "Xent::EvalMasked",                   # :76-164
        "Xent::Report",                       # :293-307
    ]),
}


def find_definition(text, qname):
    """Return (start, end, first_line_no) of the out-of-class definition `... qname(` in text, including
    the preceding `template<...>` line(s)."""
    pat = re.compile(r"(?m)^[^\n/]*\b" + re.escape(qname) + r"\s*\(")
    for m in pat.finditer(text):
        line_start = text.rfind("\n", 0, m.start()) + 1
        # the signature must be followed by a body before the next ';' at depth 0
        i = m.end()
        depth = 1
        while depth:  # close the parameter list
            c = text[i]
            depth += (c == "(") - (c == ")")
            i += 1
        j = i
        while text[j] in " \t\r\nconst":
            j += 1
        if text[j] != "{":
            continue  # a declaration, a call or an explicit instantiation
        # include template<> lines (and a return type on its own line) directly above
        start = line_start
        while True:
            prev_end = start - 1
            prev_start = text.rfind("\n", 0, prev_end) + 1
            prev = text[prev_start:prev_end].strip()
            if prev.startswith("template") or prev in ("void", "std::string", "Real"):
                start = prev_start
            else:
                break
        # brace-match the body (the extracted functions hold no braces in strings or comments)
        depth = 0
        k = j
        while True:
            c = text[k]
            depth += (c == "{") - (c == "}")
            k += 1
            if depth == 0:
                break
        return start, k, text.count("\n", 0, start) + 1
    raise SystemExit("extract_ref_ops: definition of %s not found" % qname)


def main():
    ref_root, out_dir = sys.argv[1], sys.argv[2]
    os.makedirs(out_dir, exist_ok=True)
    for group, (rel, names) in WANTED.items():
        path = os.path.join(ref_root, rel)
        text = open(path).read()
        parts = ["// GENERATED by oracle/ref_build/extract_ref_ops.py from %s -- verbatim reference code, build output, "
                 "never committed.\n" % path]
        for qn in names:
            s, e, line = find_definition(text, qn)
            parts.append("// ---- %s:%d  %s\n#line %d \"%s\"\n%s\n" % (rel, line, qn, line, path, text[s:e]))
        with open(os.path.join(out_dir, group + ".inc"), "w") as f:
            f.write("\n".join(parts))
    print("extracted:", ", ".join("%s (%d)" % (g, len(v[1])) for g, v in WANTED.items()))


if __name__ == "__main__":
    main()
