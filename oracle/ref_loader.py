"""
TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference (PimLb/POMDP_PBVI_Exploration).

Imports /root/reference/src/{mdp,pomdp}.py as they lie (read-only mount) with a stubbed
matplotlib (the reference imports it at module top: src/mdp.py:3-4, src/pomdp.py:3-6) and with
cupy absent, so every `xp` in the reference resolves to NumPy (src/pomdp.py:15-21).

/root/reference only exists in the authoring container.  Nothing that runs on the GPU box
(`pytest -m gpu`, `smoke()`, `bench.py`) may import this module; it is used by
`tests/golden/make_golden.py` (fixture generation) and by CPU tests that are skipped when the
reference mount is absent.
"""
import contextlib
import io
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("PBVI_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "pomdp.py"))


def _install_matplotlib_stub():
    if "matplotlib" in sys.modules:
        return
    names = ["matplotlib", "matplotlib.pyplot", "matplotlib.animation", "matplotlib.cm", "matplotlib.colors",
             "matplotlib.ticker", "matplotlib.patches", "matplotlib.lines"]
    mods = {n: types.ModuleType(n) for n in names}
    for n, m in mods.items():
        sys.modules[n] = m
        if "." in n:
            setattr(mods["matplotlib"], n.split(".")[1], m)
    tab = ["blue", "orange", "green", "red", "purple", "brown", "pink", "gray", "olive", "cyan"]
    hexes = ["#1f77b4", "#ff7f0e", "#2ca02c", "#d62728", "#9467bd", "#8c564b", "#e377c2", "#7f7f7f", "#bcbd22", "#17becf"]
    mods["matplotlib.colors"].TABLEAU_COLORS = {f"tab:{n}": h for n, h in zip(tab, hexes)}
    mods["matplotlib.patches"].Rectangle = object
    mods["matplotlib.lines"].Line2D = object


_ref = None


def load_reference():
    """Returns the reference's `src.pomdp` module (with `src.mdp` reachable as `.mdp_module`)."""
    global _ref
    if _ref is not None:
        return _ref
    if not reference_available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_ROOT}")
    _install_matplotlib_stub()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with contextlib.redirect_stdout(io.StringIO()):
        import src.mdp as ref_mdp      # noqa
        import src.pomdp as ref_pomdp  # noqa
    ref_pomdp.mdp_module = ref_mdp
    _ref = ref_pomdp
    return _ref


@contextlib.contextmanager
def quiet():
    """The reference logs with print(); silence it around model construction / solve."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield
