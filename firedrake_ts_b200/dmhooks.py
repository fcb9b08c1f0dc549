"""The two dmhooks entry points the reference relies on
(firedrake_ts/solving_utils.py:245-246, firedrake_ts/ts_solver.py:266-268,
:356-363): a stack of application contexts on the DM."""
from contextlib import contextmanager


def get_appctx(dm):
    if not dm._appctx:
        raise RuntimeError("no application context attached to the DM: call inside dmhooks.add_hooks(...)")
    return dm._appctx[-1]


@contextmanager
def add_hooks(dm, solver, appctx=None, save=True):
    dm._appctx.append(appctx)
    try:
        yield
    finally:
        dm._appctx.pop()
