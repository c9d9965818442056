"""TEST INFRASTRUCTURE ONLY.  Runs individual functions of the reference's trainer VERBATIM: `main_avatar.py` cannot be
imported in this image (it pulls in the dataset / renderer stack at import time), but its methods are plain torch code —
their source text is cut out of the installed copy (baseline/_ref/AnimatableGaussians/main_avatar.py, or /root/reference
where that exists) with `ast` and compiled as-is.  Nothing is edited; the test supplies `self`."""
import ast
import os
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = (os.path.join(ROOT, "baseline", "_ref", "AnimatableGaussians"), "/root/reference")


def ref_file(rel):
    for base in CANDIDATES:
        p = os.path.join(base, rel)
        if os.path.exists(p):
            return p
    return None


def method(rel, cls, name, namespace):
    """-> the function object of `cls.name` from the reference file `rel`, compiled from its unmodified source in `namespace`."""
    path = ref_file(rel)
    if path is None:
        return None
    src = open(path).read()
    tree = ast.parse(src)
    for node in ast.walk(tree):
        if isinstance(node, ast.ClassDef) and node.name == cls:
            for f in node.body:
                if isinstance(f, ast.FunctionDef) and f.name == name:
                    f.decorator_list = []
                    code = textwrap.dedent(ast.get_source_segment(src, f))
                    if code.lstrip().startswith("@"):
                        code = code[code.index("def "):]
                    ns = dict(namespace)
                    exec(compile(code, path, "exec"), ns)
                    return ns[name]
    return None
