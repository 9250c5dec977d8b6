"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified-in-arithmetic* reference `lopq` package.

The reference (`/root/reference/lopq/lopq/{utils,model,search,eval}.py`) is Python 2.  This
loader reads those files where they lie, applies a fixed list of *token level* py2->py3 patches
(no arithmetic is touched) and execs them as the package ``ref_lopq``.  It is used only

  * by ``tests/golden/make_golden.py`` to generate the committed golden vectors, and
  * by ``tests/test_oracle_vs_reference.py`` (skipped when /root/reference is absent)
    to pin ``oracle/lopq_oracle.py`` against the real reference.

`/root/reference` does not exist on the GPU box; nothing in the product, the gpu tests, smoke()
or bench.py imports this module.

Patch list:
  xrange -> range; `print X` -> print(X); builtin reduce -> functools.reduce;
  integer `/` -> `//` at the listed integer sites only; list-returning map() wrapped in list();
  time.clock -> time.perf_counter; implicit relative imports are replaced by building the
  package by hand.
"""
import os
import re
import sys
import types

REF_DIR = os.environ.get("LOPQ_REFERENCE_DIR", "/root/reference/lopq/lopq")

# (file, 1-based line numbers) whose ` / ` is a py2 *integer* division
_INT_DIV_LINES = {
    "utils.py": {19, 86, 170},
    "model.py": {41, 407, 408, 433, 434, 491, 744, 797, 873},
}
# lines whose map(...) result is consumed as a list
_LIST_MAP_LINES = {
    "utils.py": {29},
    "search.py": {219, 222},
    "model.py": {726, 740, 741, 742, 744, 810, 812},
}

_PRINT_RE = re.compile(r"^(\s*)print (.+?)\s*$")


def available():
    return os.path.isfile(os.path.join(REF_DIR, "model.py"))


def _patch(fname, src):
    out = []
    for ln, line in enumerate(src.split("\n"), 1):
        if ln in _INT_DIV_LINES.get(fname, ()):
            line = line.replace(" / ", " // ")
        if ln in _LIST_MAP_LINES.get(fname, ()):
            # wrap the outermost map( ... ) of an assignment in list( ... )
            line = re.sub(r"= map\((.*)\)\s*$", r"= list(map(\1))", line)
        m = _PRINT_RE.match(line)
        if m and not line.lstrip().startswith("#"):
            line = "%sprint(%s)" % (m.group(1), m.group(2))
        line = line.replace("xrange(", "range(")
        line = line.replace("time.clock()", "time.perf_counter()")
        out.append(line)
    return "from functools import reduce\n" + "\n".join(out)


def load(name="ref_lopq"):
    """Return the patched reference package (cached in sys.modules)."""
    if name in sys.modules:
        return sys.modules[name]
    if not available():
        raise ImportError("reference lopq sources not found under %s" % REF_DIR)
    pkg = types.ModuleType(name)
    pkg.__path__ = []  # mark as package
    pkg.__package__ = name
    sys.modules[name] = pkg
    for mod in ("utils", "model", "search", "eval"):
        fname = mod + ".py"
        with open(os.path.join(REF_DIR, fname)) as f:
            src = _patch(fname, f.read())
        m = types.ModuleType("%s.%s" % (name, mod))
        m.__package__ = name
        m.__file__ = os.path.join(REF_DIR, fname)
        sys.modules[m.__name__] = m
        # "from functools import reduce" shifts line numbers by one in tracebacks only
        exec(compile(src, m.__file__, "exec"), m.__dict__)
        setattr(pkg, mod, m)
    pkg.LOPQModel = pkg.model.LOPQModel
    pkg.LOPQModelPCA = pkg.model.LOPQModelPCA
    pkg.LOPQCode = pkg.model.LOPQCode
    pkg.LOPQSearcher = pkg.search.LOPQSearcher
    pkg.multisequence = pkg.search.multisequence
    return pkg
