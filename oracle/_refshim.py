"""Import the real reference's metric functions (build container only).

``/root/reference`` is a plain directory, not an installed package, and its
modules import each other as ``LDiffusion.<mod>``.  This shim aliases the
directory as package ``LDiffusion``, stubs the unused ``tifffile`` import
(``utils.py:1``) and puts the vendored nnU-Net on ``sys.path`` so that
``nnunetv2.paths`` (``utils.py:12``) resolves.

The reference does not exist on the GPU box: callers must check
``available()`` first.  Only ``tests/golden/make_golden.py`` and the
"reference present" tests use this.
"""
import importlib.machinery
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("LDIFF_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "utils.py")) and os.path.isfile(
        os.path.join(REF_ROOT, "evaluate.py")
    )


def load():
    """Returns (utils_module, evaluate_module) of the reference."""
    if not available():
        raise RuntimeError(f"reference checkout not found at {REF_ROOT}")
    if "LDiffusion" not in sys.modules:
        spec = importlib.machinery.ModuleSpec("LDiffusion", None, is_package=True)
        spec.submodule_search_locations = [REF_ROOT]
        pkg = importlib.util.module_from_spec(spec)
        pkg.__path__ = [REF_ROOT]
        sys.modules["LDiffusion"] = pkg
    sys.modules.setdefault("tifffile", types.ModuleType("tifffile"))
    model_dir = os.path.join(REF_ROOT, "model")
    if model_dir not in sys.path:
        sys.path.insert(0, model_dir)
    # nnunetv2.paths prints three warnings about unset env vars; silence them.
    devnull = open(os.devnull, "w")
    old = sys.stdout
    sys.stdout = devnull
    try:
        import LDiffusion.utils as ref_utils
        import LDiffusion.evaluate as ref_evaluate
    finally:
        sys.stdout = old
        devnull.close()
    return ref_utils, ref_evaluate


def load_functions(path, names):
    """Compile only the named top-level functions of a reference module whose imports cannot be satisfied
    here (e.g. nnU-Net's evaluator pulls in batchgenerators and SimpleITK) — the functions' own source, run
    as is, with numpy and the typing names they use in scope."""
    import ast
    from typing import List, Tuple, Union

    import numpy as np
    tree = ast.parse(open(path).read())
    ns = {"np": np, "List": List, "Tuple": Tuple, "Union": Union}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module([node], []), path, "exec"), ns)
    return [ns[n] for n in names]
