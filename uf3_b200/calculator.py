"""`UFCalculator` — the inference-path entry point (ASE `Calculator` protocol), computed by
the CUDA path.

API mirror of `/root/reference/uf3/forcefield/calculator.py:40-153,399-404,490-520`:
`UFCalculator(model)` with `implemented_properties = ['energy', 'forces', 'stress']`,
`calculate(atoms, properties, system_changes)` filling `self.results`, the attributes
`solutions` / `pair_potentials` / `trio_potentials` (coefficient vectors and decompressed
coefficient grids per interaction — the reference stores `ndsplines.NDSpline` objects
there; here they are the arrays the kernels contract against).  Stress: the reference takes
central finite differences of the energy over strain (ASE's `calculate_numerical_stress`,
12 energy evaluations); here the evaluator kernel accumulates the analytic strain derivative
in the same pair / triangle walk (`uf3b_energy_forces(..., virial)`), and
`UFCalculator(model, numerical_stress=True)` selects the reference's method.  If ASE is
installed the class derives from
`ase.calculators.calculator.Calculator`; otherwise from a minimal stand-in with the
same protocol (`get_potential_energy`, `get_forces`, `get_stress`, `results`).
"""
import numpy as np

from uf3_b200 import geometry
from uf3_b200.atoms import frame_arrays

try:  # pragma: no cover - ASE is optional
    from ase.calculators.calculator import Calculator as _Base, all_changes
except ImportError:
    all_changes = ["positions", "numbers", "cell", "pbc", "initial_charges", "initial_magmoms"]

    class _Base:
        """The part of ASE's Calculator protocol that UF3 relies on."""
        implemented_properties = []

        def __init__(self, **kwargs):
            self.results = {}
            self.atoms = None
            self.parameters = dict(kwargs)

        def calculate(self, atoms=None, properties=None, system_changes=all_changes):
            if atoms is not None:
                self.atoms = atoms.copy()

        def get_property(self, name, atoms=None, allow_calculation=True):
            if name not in self.implemented_properties:
                raise NotImplementedError(f"{name} property not implemented")
            self.calculate(atoms, [name], all_changes)
            return self.results[name]

        def get_potential_energy(self, atoms=None, force_consistent=False):
            return self.get_property("energy", atoms)

        def get_forces(self, atoms=None):
            return self.get_property("forces", atoms)

        def get_stress(self, atoms=None):
            return self.get_property("stress", atoms)

        def calculate_numerical_stress(self, atoms, d=1e-6, voigt=True):
            """Central differences of the energy over the six strain components."""
            stress = np.zeros((3, 3))
            cell = atoms.get_cell()
            volume = atoms.get_volume()
            for i in range(3):
                for j in range(i, 3):
                    energies = []
                    for sign in (1, -1):
                        strain = np.eye(3)
                        if i == j:
                            strain[i, i] += sign * d
                        else:
                            strain[i, j] += sign * d / 2
                            strain[j, i] += sign * d / 2
                        trial = atoms.copy()
                        trial.set_cell(np.dot(cell, strain), scale_atoms=True)
                        energies.append(self.get_potential_energy(trial, force_consistent=True))
                    stress[i, j] = stress[j, i] = (energies[0] - energies[1]) / (2 * d * volume)
            if voigt:
                return stress.flat[[0, 4, 8, 5, 2, 1]]
            return stress


def coefficients_by_interaction(element_list, interactions_map, partition_sizes, coefficients):
    """Flat coefficient vector -> {element or interaction: slice} (calculator.py:490-520)."""
    pieces = np.array_split(coefficients, np.cumsum(partition_sizes)[:-1])
    solutions = dict(zip(element_list, pieces[:len(element_list)]))
    keys = list(interactions_map[2]) + list(interactions_map.get(3, []))
    for idx, key in enumerate(keys):
        if len(element_list) + idx < len(pieces):
            solutions[key] = pieces[len(element_list) + idx]
    return solutions


class UFCalculator(_Base):
    implemented_properties = ["energy", "forces", "stress"]

    def __init__(self, model, device=None, numerical_stress=False, **kwargs):
        super().__init__(**kwargs)
        self.numerical_stress = numerical_stress
        self.bspline_config = model.bspline_config
        self.model = model
        self.device = device
        self.solutions = coefficients_by_interaction(self.element_list, self.interactions_map,
                                                     self.partition_sizes, model.coefficients)
        self.pair_potentials = {pair: (self.bspline_config.knots_map[pair], self.solutions[pair])
                                for pair in self.interactions_map[2]}
        if self.degree > 2:
            self.trio_potentials = {
                trio: (self.bspline_config.knots_map[trio],
                       self.bspline_config.decompress_3B(self.solutions[trio], trio))
                for trio in self.interactions_map[3]}
        self._engine = None

    def __getstate__(self):
        state = dict(self.__dict__)
        state["_engine"] = None
        return state

    @property
    def engine(self):
        if self._engine is None:
            from uf3_b200.engine import Engine
            self._engine = Engine(self.bspline_config, device=self.device)
            self._engine.set_coefficients(self.model.coefficients)
        return self._engine

    def __repr__(self):
        return "\n".join(["UFCalculator:", repr(self.model)])

    degree = property(lambda self: self.bspline_config.degree)
    element_list = property(lambda self: self.bspline_config.element_list)
    interactions_map = property(lambda self: self.bspline_config.interactions_map)
    r_min_map = property(lambda self: self.bspline_config.r_min_map)
    r_max_map = property(lambda self: self.bspline_config.r_max_map)
    r_cut = property(lambda self: self.bspline_config.r_cut)
    partition_sizes = property(lambda self: self.bspline_config.partition_sizes)
    coefficients = property(lambda self: self.model.coefficients)
    chemical_system = property(lambda self: self.bspline_config.chemical_system)

    # ------------------------------------------------------------------ ASE protocol
    def calculate(self, atoms=None, properties=None, system_changes=tuple(all_changes)):
        if properties is None:
            properties = self.implemented_properties
        super().calculate(atoms, properties, system_changes)
        want_e = "energy" in properties or "free_energy" in properties
        want_f = "forces" in properties
        if want_e or want_f:
            energy, forces = self._evaluate(atoms, want_e, want_f)
            if want_e:
                self.results["energy"] = energy
                self.results["free_energy"] = energy
            if want_f:
                self.results["forces"] = forces
        if "stress" in properties:
            self.results["stress"] = self._get_stress(atoms)

    def _evaluate(self, atoms, want_e, want_f):
        positions, numbers, cell, pbc = frame_arrays(atoms)
        images = geometry.image_table(cell, pbc, self.r_cut) if np.any(pbc) else None
        eng = self.engine
        eng.build_neighbors(positions, numbers, images=images)
        return eng.energy_forces(energy=want_e, forces=want_f)

    def _get_potential_energy(self, atoms=None, force_consistent=None):
        energy, _ = self._evaluate(atoms, True, False)
        if force_consistent is True:   # calculator.py:166-169: drop the one-body offsets
            energy -= float(sum(self.solutions[el][0] * n for el, n in zip(
                self.element_list, self.chemical_system.get_composition_tuple(atoms))))
        return energy

    def _get_forces(self, atoms=None):
        return self._evaluate(atoms, False, True)[1]

    def _get_stress(self, atoms=None, **kwargs):
        """Voigt stress [xx, yy, zz, yz, xz, xy] (calculator.py:399-404)."""
        if self.numerical_stress or kwargs:
            return self.calculate_numerical_stress(atoms, **kwargs)
        positions, numbers, cell, pbc = frame_arrays(atoms)
        images = geometry.image_table(cell, pbc, self.r_cut) if np.any(pbc) else None
        eng = self.engine
        eng.build_neighbors(positions, numbers, images=images)
        _, _, w = eng.energy_forces(energy=False, forces=False, virial=True)
        volume = abs(np.linalg.det(np.asarray(cell, dtype=np.float64)))
        return (w / volume).flat[[0, 4, 8, 5, 2, 1]]

    knot_subintervals = property(lambda self: self.bspline_config.knot_subintervals)

    def calculation_required(self, atoms, quantities):
        """ASE's legacy protocol (calculator.py:438-447): everything but energy / forces / stress is not
        available; otherwise a calculation is needed when the atoms changed."""
        if any(q in quantities for q in ("magmom", "charges")):
            return True
        return self.atoms is None or self.atoms != atoms or any(q not in self.results for q in quantities)

    # ------------------------------------------------------------------ relaxation
    def relax_fmax(self, geom, fmax=0.05, relax_cell=True, verbose=False, timeout=60.0, max_steps=2000,
                   **kwargs):
        """Minimise the largest force (calculator.py:406-436).  With ASE installed this is the
        reference's recipe (BFGSLineSearch on an ExpCellFilter when the cell is periodic and
        `relax_cell`).  Without ASE the same degrees of freedom — atoms in the frame of the initial
        cell plus the deformation gradient, scaled by the atom count as ASE's cell filters do — are
        relaxed with FIRE; the cell gradient comes from the ANALYTIC virial of the evaluator kernel,
        one launch per step, where the reference takes twelve strained-cell energies.  Returns the
        relaxed copy of `geom`; warns and returns the current state after `timeout` seconds."""
        import time
        import warnings
        geom = geom.copy()
        periodic_cell = bool(np.all(geom.get_pbc())) and relax_cell
        try:  # pragma: no cover - ASE is optional
            from ase import constraints as ase_constraints, optimize as ase_optim
            import os
            geom.calc = self
            target = ase_constraints.ExpCellFilter(geom) if periodic_cell else geom
            optimizer = ase_optim.BFGSLineSearch(target, logfile="-" if verbose else os.devnull, **kwargs)
            t0 = time.time()
            for _ in optimizer.irun(fmax=fmax):
                if time.time() - t0 > timeout:
                    warnings.warn("Relaxation timed out.", RuntimeWarning)
                    break
            return geom
        except ImportError:
            pass

        positions, numbers, cell0, pbc = frame_arrays(geom)
        cell0 = np.array(cell0, dtype=np.float64)
        n = len(positions)
        ref = np.array(positions, dtype=np.float64)          # atoms in the frame of the initial cell
        deform = np.eye(3)
        scale = float(n)                                     # ASE's cell_factor
        eng = self.engine

        def gradients():
            pos = ref @ deform.T
            cell = cell0 @ deform.T
            images = geometry.image_table(cell, pbc, self.r_cut) if np.any(pbc) else None
            eng.build_neighbors(pos, numbers, images=images)
            if periodic_cell:
                energy, forces, virial = eng.energy_forces(energy=True, forces=True, virial=True)
                g_cell = -(virial @ np.linalg.inv(deform).T) / scale
            else:
                energy, forces = eng.energy_forces(energy=True, forces=True)
                g_cell = np.zeros((3, 3))
            return energy, np.vstack([forces @ deform, g_cell]), pos, cell

        # FIRE (Bitzek et al. 2006) with ASE's default parameters
        dt, dt_max, n_min, f_inc, f_dec, a_start, f_a, max_move = 0.1, 1.0, 5, 1.1, 0.5, 0.1, 0.99, 0.2
        alpha, uphill_free = a_start, 0
        velocity = np.zeros((n + 3, 3))
        t0 = time.time()
        for step in range(max_steps):
            energy, force, pos, cell = gradients()
            residual = float(np.sqrt((force ** 2).sum(axis=1).max()))
            if verbose:
                print(f"relax_fmax step {step:4d}  E = {energy:.8f}  fmax = {residual:.6f}")
            if residual < fmax:
                break
            if time.time() - t0 > timeout:
                warnings.warn("Relaxation timed out.", RuntimeWarning)
                break
            power = float(np.vdot(force, velocity))
            if power > 0.0:
                f_norm = np.sqrt(np.vdot(force, force))
                velocity = (1.0 - alpha) * velocity + alpha * force / f_norm * np.sqrt(np.vdot(velocity, velocity))
                if uphill_free > n_min:
                    dt = min(dt * f_inc, dt_max)
                    alpha *= f_a
                uphill_free += 1
            else:
                velocity[:] = 0.0
                alpha, dt, uphill_free = a_start, dt * f_dec, 0
            velocity += dt * force
            move = dt * velocity
            longest = float(np.sqrt(np.vdot(move, move)))
            if longest > max_move:
                move *= max_move / longest
            ref += move[:n]
            deform += move[n:] / scale
        else:
            warnings.warn("Relaxation stopped after max_steps.", RuntimeWarning)
        geom.set_cell(cell0 @ deform.T, scale_atoms=False)
        geom.set_positions(ref @ deform.T)
        return geom
