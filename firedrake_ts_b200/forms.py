"""Form catalogue: the weak forms of the reference's examples, as descriptors.

The reference hands UFL forms to Firedrake/TSFC, which JIT-compiles an element
kernel per form.  This engine has no code generator:
it ships hand-written kernels for the *form families* that the reference's
``examples/`` exercise and a descriptor object
that names the family, its constants, and the ``u`` / ``udot`` Functions:

  heat           examples/heat.py:11          F = (u_t, v) + (grad u, grad v) - (1, v)
  burgers        examples/burgers.py:21-25    F = (u_t, v) + ((u.grad)u, v) + nu (grad u, grad v)
  bbm            examples/bbm.py:35-40        F = (u_t, v) + (u_x, v) + (u u_x, v) + (u_t_x, v_x)
  cahn_hilliard  examples/cahn-hilliard.py:26-38
                 F = (c_t, q) + (grad mu, grad q) + (mu, v) - (f'(c), v) - lambda (grad c, grad v)
                 f = 100 c^2 (1-c)^2
  diffusion      examples/heat-explicit.py:13-14   a (u_t, v) + k (grad u, grad v) - s (1, v)
                 the two sides of the IMEX split F(u_t, u, t) = G(u, t):
                 F = (u_t, v) is (a, k, s) = (1, 0, 0);  G = -((grad u, grad v) - (1, v)) is (0, -1, -1)

J is always ``shift * dF/dudot + dF/du`` (firedrake_ts/ts_solver.py:108); the
derivative is taken by hand in the kernels (and in the oracle), ``shift`` stays
a run-time scalar exactly like the reference's ``Constant``.
"""
from __future__ import annotations

from dataclasses import dataclass, field

HEAT, BURGERS, BBM, CAHN_HILLIARD, DIFFUSION = 0, 1, 2, 3, 4
FAMILY_NAMES = {HEAT: "heat", BURGERS: "burgers", BBM: "bbm", CAHN_HILLIARD: "cahn_hilliard",
                DIFFUSION: "diffusion"}


def quadrature_degree(family: int, degree: int) -> int:
    """UFL's estimated total polynomial degree for each family."""
    k = degree
    if family in (HEAT, DIFFUSION):
        return 2 * k
    if family in (BURGERS, BBM):
        return max(3 * k - 1, 2 * k)
    if family == CAHN_HILLIARD:
        return 4 * k
    raise ValueError(family)


@dataclass
class Form:
    """A residual form F(u, udot; v) of one of the supported families."""
    family: int
    u: object
    udot: object
    params: tuple = (0.0, 0.0, 0.0, 0.0)
    rank: int = 1  # 1 = residual (linear form), 2 = Jacobian (bilinear form)
    shift: object = None

    @property
    def space(self):
        return self.u.function_space()

    @property
    def name(self):
        return FAMILY_NAMES[self.family]

    def arguments(self):
        """mimics ufl.Form.arguments(): one entry per test/trial function."""
        return (self.space,) * self.rank

    @property
    def qdegree(self):
        return quadrature_degree(self.family, self.space.degree)


def heat(u, udot, source: float = 1.0) -> Form:
    assert u.function_space().ncomp == 1
    return Form(HEAT, u, udot, (float(source), 0.0, 0.0, 0.0))


def diffusion(u, udot, mass: float = 1.0, diffusivity: float = 1.0, source: float = 0.0) -> Form:
    """a (u_t, v) + k (grad u, grad v) - s (1, v); params = (s, a, k)."""
    assert u.function_space().ncomp == 1
    return Form(DIFFUSION, u, udot, (float(source), float(mass), float(diffusivity), 0.0))


def heat_explicit(u, udot, source: float = 1.0):
    """The IMEX split of examples/heat-explicit.py:13-14: returns (F, G) with
    F = (u_t, v) and G = -((grad u, grad v) - (source, v))."""
    return (diffusion(u, udot, mass=1.0, diffusivity=0.0, source=0.0),
            diffusion(u, udot, mass=0.0, diffusivity=-1.0, source=-float(source)))


def mass_form(F: Form) -> Form:
    """derivative(F, udot) as a form whose IJacobian at shift = 1 is the mass matrix (used by the
    right-hand-side projection solver, reference solving_utils.py:429-445)."""
    if F.family == HEAT:
        return diffusion(F.u, F.udot, mass=1.0, diffusivity=0.0)
    if F.family == DIFFUSION:
        return diffusion(F.u, F.udot, mass=F.params[1], diffusivity=0.0)
    raise NotImplementedError(f"mass matrix of the {F.name} family is not in the catalogue")


def burgers(u, udot, nu: float = 1.0e-4) -> Form:
    V = u.function_space()
    assert V.ncomp == V.dim and V.layout == "interleaved", "Burgers needs a VectorFunctionSpace"
    return Form(BURGERS, u, udot, (float(nu), 0.0, 0.0, 0.0))


def bbm(u, udot) -> Form:
    assert u.function_space().ncomp == 1
    return Form(BBM, u, udot)


def cahn_hilliard(u, udot, lmbda: float = 1.0e-2) -> Form:
    V = u.function_space()
    assert V.ncomp == 2 and V.layout == "blocked", "Cahn-Hilliard needs the mixed space V*V"
    return Form(CAHN_HILLIARD, u, udot, (float(lmbda), 0.0, 0.0, 0.0))


def derivative(form: Form, shift) -> Form:
    """J = shift*dF/dudot + dF/du for a catalogue form (ts_solver.py:108)."""
    return Form(form.family, form.u, form.udot, form.params, rank=2, shift=shift)
