"""shared builders for the tests: the synthetic configurations of SURVEY.md section 8(d)."""
import numpy as np
import torch

from firedrake_ts_b200.workloads import make_problem, rel_err  # noqa: F401


class OracleAssembler:
    """plugs the CPU oracle behind _TSContext's assembler interface (tests only): reference
    trajectories are produced through the very same callbacks and time stepper."""

    def __init__(self, form, bcs):
        from oracle import Oracle

        self.o = Oracle.from_form(form, bcs)
        self.device = torch.device("cpu")

    def pattern(self):
        return self.o.pattern()

    def ifunction(self, t, x, xdot, out=None):
        F = torch.from_numpy(self.o.ifunction(t, x, xdot))
        if out is not None:
            out.copy_(F)
            return out
        return F

    def ijacobian(self, t, x, xdot, shift, out=None):
        J = torch.from_numpy(self.o.ijacobian(t, x, xdot, shift))
        if out is not None:
            out.copy_(J)
            return out
        return J
