"""TEST INFRASTRUCTURE (checker; never imported by the product): numpy restatement of the reference's order
parameters, /root/reference/nematic.py -- nematic_order_param :83-137 (minimum-image end-to-end vector,
Q = mean(1.5 u u^T - 0.5 I), largest eigenvalue) and trans_order_param :30-55 with centre_of_masses :57-81
(fractional mean over beads a0:a1, i.e. WITHOUT the last bead; the bead itself for Nbeads = 1).  The reference's
trans_order_param indexes r[i] for i < Nchains*Nbeads although r has Nchains entries (IndexError for Nbeads > 1);
the restatement sums over the Nchains centres, which is what it computes when it runs (Nbeads = 1).
Parity: pinned by executing the unmodified /root/reference/nematic.py with an ase stand-in on the same configurations
(tests/test_reference_live_cpu.py::test_reference_nematic_functions_match_the_numpy_restatement, 1e-12)."""
import numpy as np


def mic_vector(d, cell):
    """ase's get_distance(mic=True, vector=True) for vectors shorter than half the cell width."""
    s = d @ np.linalg.inv(cell)
    s -= np.rint(s)
    return s @ cell


def nematic_order_param(cells, positions, Nbeads, Nchains):
    out = []
    for cell, pos in zip(cells, positions):
        if Nbeads == 1:
            out.append(0.0)
            continue
        Q = np.zeros((3, 3))
        for i in range(Nchains):
            a0, a1 = i * Nbeads, i * Nbeads + Nbeads - 1
            v = mic_vector(pos[a1] - pos[a0], cell)
            v = v / np.linalg.norm(v)
            Q += 1.5 * np.outer(v, v) - 0.5 * np.identity(3)
        out.append(np.amax(np.linalg.eigvalsh(Q / Nchains)))
    return np.array(out)


def trans_order_param(cells, positions, Nbeads, Nchains):
    out = []
    for cell, pos in zip(cells, positions):
        frac = (pos @ np.linalg.inv(cell)) % 1.0   # ase get_scaled_positions() wraps
        rho = 0.0 + 0.0j
        for i in range(Nchains):
            a0, a1 = i * Nbeads, i * Nbeads + Nbeads - 1
            r = frac[a0] if Nbeads == 1 else frac[a0:a1].mean(0)
            rho += np.exp(1j * 2 * np.pi * np.sum(r))
        out.append(abs(rho / Nchains))
    return np.array(out)
