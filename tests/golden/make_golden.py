"""Extract the golden vectors / known-answer values that chiron's own tests hold for the particle hot
path and write them to tests/golden/reference_goldens.json.

The reference cannot be imported here (jax, openmm, openmmtools are not installed), so the goldens
are the numeric literals of its test-suite, pulled out of the source with `ast` (no reference code
is copied, only the asserted numbers).  Run in the build container, where /root/reference exists:

    python tests/golden/make_golden.py
"""
import ast
import json
import os
import sys

REF = os.environ.get("CHIRON_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_goldens.json")


def _num(node):
    if isinstance(node, ast.Constant) and isinstance(node.value, (int, float)):
        return node.value
    if isinstance(node, ast.UnaryOp) and isinstance(node.op, ast.USub):
        return -_num(node.operand)
    raise ValueError


def _nested(node):
    if isinstance(node, (ast.List, ast.Tuple)):
        return [_nested(e) for e in node.elts]
    return _num(node)


def numeric_lists(func_node, min_len=2):
    """All outermost list literals made only of numbers (nested allowed) inside a function."""
    found = []

    def visit(node):
        if isinstance(node, ast.List):
            try:
                val = _nested(node)
                if len(val) >= min_len:
                    found.append((node.lineno, val))
                return
            except ValueError:
                pass
        for child in ast.iter_child_nodes(node):
            visit(child)

    visit(func_node)
    return found


def functions(path):
    with open(path) as fh:
        tree = ast.parse(fh.read())
    return {n.name: n for n in tree.body if isinstance(n, ast.FunctionDef)}


def main():
    t = os.path.join(REF, "chiron", "tests")
    out = {"source": "choderalab/chiron chiron/tests (numeric literals only)", "entries": {}}
    e = out["entries"]

    f = functions(os.path.join(t, "test_mcmc.py"))
    lists = numeric_lists(f["test_sample_from_harmonic_osciallator"], 5)
    line, vals = [(l, v) for l, v in lists if len(v) == 5][0]
    e["langevin_ho_energy_trace"] = {
        "cite": f"chiron/tests/test_mcmc.py:{line}", "values": vals,
        "setup": "HarmonicOscillator() K=100 kcal/mol/A^2, m=39.948, x0=0; PRNG seed 1234, first key; "
                 "LangevinIntegrator(dt=2 fs, gamma=1/ps, refresh_velocities=True, report_interval=1), "
                 "T=300 K, 5 steps; potential energy after each step; jnp.allclose defaults"}
    e["mc_barostat_counts"] = {
        "cite": "chiron/tests/test_mcmc.py:451-452", "n_proposed": 10, "n_accepted": 8,
        "setup": "8 particles on the unit cube corners (nm), box 10 nm, IdealGasPotential, T=300 K, "
                 "P=1 atm, MonteCarloBarostatMove(volume_max_scale=0.1, number_of_moves=10), seed 1234"}

    f = functions(os.path.join(t, "test_utils.py"))
    lists = numeric_lists(f["test_reporter"], 20)
    line, vals = [(l, v) for l, v in lists if len(v) == 20][0]
    e["langevin_ho_energy_trace_20"] = {
        "cite": f"chiron/tests/test_utils.py:{line}", "values": vals,
        "setup": "HarmonicOscillatorPotential default k=1 kcal/mol/A^2, x0=0, m=39.948; seed 1234 first key; "
                 "LangevinIntegrator(report_interval=1) dt=1 fs, gamma=1/ps, T=300 K, 20 steps "
                 "(the reference passes this list to np.allclose without assert)"}

    f = functions(os.path.join(t, "test_pairs.py"))
    lists = numeric_lists(f["test_neighborlist_pair_multiple_particles"], 8)
    e["nlist_cube8"] = {
        "cite": "chiron/tests/test_pairs.py:240-318",
        "n_neighbors": [v for _, v in lists if v == [7, 6, 5, 4, 3, 2, 1, 0]][0],
        "n_interacting_cutoff_1p1": [v for _, v in lists if len(v) == 8 and v[0] == 3][0],
        "neighbor_list_8x17": [v for _, v in lists if len(v) == 8 and isinstance(v[0], list) and len(v[0]) == 17][0],
        "setup": "2x2x2 grid spacing 1 nm, box 10 nm; (cutoff 2.1, skin 0.1) and (cutoff 1.1, skin 1.1), "
                 "n_max_neighbors=5 grows to 17"}
    lists = numeric_lists(f["test_neighborlist_pair"], 2)
    e["nlist_pair2"] = {
        "cite": "chiron/tests/test_pairs.py:60-160",
        "neighbor_list": [[1, 1, 1, 1, 1], [0, 0, 0, 0, 0]],
        "neighbor_mask": [[1, 0, 0, 0, 0], [0, 0, 0, 0, 0]],
        "n_neighbors": [1, 0],
        "literal_lists_found": len(lists),
        "setup": "particles (0,0,0),(1,0,0) nm, box 10, cutoff 1.1, skin 0.1, n_max_neighbors=5; "
                 "dist all 1.0, r_ij rows -1/+1; check false, true after +0.1 shift, true after N change"}
    lists = numeric_lists(f["test_pair_list_multiple_particles"], 8)
    e["pairlist_cube8"] = {
        "cite": "chiron/tests/test_pairs.py:415-487",
        "all_pairs": [v for _, v in lists if len(v) == 8 and isinstance(v[0], list) and len(v[0]) == 7
                      and isinstance(v[0][0], int)][0],
        "distances": [v for _, v in lists if len(v) == 8 and isinstance(v[0], list) and len(v[0]) == 7
                      and isinstance(v[0][0], float)][0]}
    lists = numeric_lists(f["test_orthogonal_periodic_displacement"], 2)
    e["space_periodic"] = {
        "cite": "chiron/tests/test_pairs.py:14-45",
        "p1": [[0, 0, 0], [0, 0, 0]], "p2": [[1, 0, 0], [6, 0, 0]], "box": 10.0,
        "r_ij": [[-1.0, 0.0, 0.0], [4.0, 0.0, 0.0]], "dist": [1, 4],
        "wrap_in": [[11, 0, 0], [-1, 0, 0], [5, 0, 0], [5, 12, -1]],
        "wrap_out": [[1, 0, 0], [9, 0, 0], [5, 0, 0], [5, 2, 9]],
        "literal_lists_found": len(lists)}

    f = functions(os.path.join(t, "test_potential.py"))
    lists = numeric_lists(f["test_harmonic_oscillator_potential"], 5)
    e["ho_energies"] = {
        "cite": "chiron/tests/test_potential.py:59-93",
        "k_kcal_per_mol_A2": 100.0,
        "positions_A": [[0.0, 0.0, 0.0], [0.2, 0.2, 0.2], [0.2, 0.0, 0.0], [-0.2, 0.0, 0.0], [-0.0, 0.2, 0.0]],
        "energies": [v for _, v in lists if len(v) == 5 and not isinstance(v[0], list)][0]}
    e["lj_two_particles"] = {
        "cite": "chiron/tests/test_potential.py:155-230",
        "setup": "sigma=1 nm, epsilon=1 kJ/mol, cutoff 3, skin 0.5, box 10; particle 2 at x = i*0.25*2^(1/6), "
                 "i=1..10; energy isclose analytic, force allclose(atol=1e-5) analytic"}
    with open(OUT, "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit(f"{REF} not found: the goldens can only be regenerated where the reference is mounted")
    main()
