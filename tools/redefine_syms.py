#!/usr/bin/env python
"""Emit `objcopy --redefine-sym old=new` arguments that rename C++ functions inside ONE object file.

    objcopy $(python tools/redefine_syms.py prover.o cpu_reference_ waffle::Prover::construct_proof waffle::Prover::reset) prover.o prover.renamed.o

Every symbol of the object (defined or referenced) whose demangled name is `<qualified name>(...)` for one of the
given fully qualified function names gets its last name component prefixed: the symbol table entry is rewritten, not
the source text, so unrelated tokens that merely spell the same word (members, locals, other classes' methods) stay
untouched — the robust form of the `-Dname=prefix_name` recipe.  Used by tests/cpp/Makefile and INTEGRATION.md to keep
the reference's CPU bodies linkable next to the GPU shims that take over their names.
"""
import subprocess
import sys


def main():
    obj, prefix, names = sys.argv[1], sys.argv[2], sys.argv[3:]
    syms = [l.split()[-1] for l in subprocess.check_output(["nm", obj], text=True).splitlines() if l.strip()]
    syms = sorted(set(s for s in syms if s.startswith("_Z")))
    if not syms:
        return
    dem = subprocess.check_output(["c++filt"], input="\n".join(syms) + "\n", text=True).splitlines()
    out = []
    for s, d in zip(syms, dem):
        for q in names:
            if not d.startswith(q + "("):
                continue
            last = q.split("::")[-1]
            old = "%d%s" % (len(last), last)
            new = "%d%s%s" % (len(prefix) + len(last), prefix, last)
            # the function's own name is the last <length><identifier> of the nested name, i.e. the one followed by 'E'
            idx = s.find(old + "E")
            if idx < 0:
                raise SystemExit("redefine_syms: cannot locate %s in %s" % (old, s))
            out.append("--redefine-sym %s=%s" % (s, s[:idx] + new + s[idx + len(old):]))
    found = set(o.split("=")[0] for o in out)
    if not out:
        raise SystemExit("redefine_syms: none of %s found in %s" % (names, obj))
    print(" ".join(out))


if __name__ == "__main__":
    main()
