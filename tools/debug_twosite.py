"""Debug helper (GPU box): run the two-site TDVP fixture through the oracle and the device path
in lockstep, recording the singular values seen by every split."""
import os, sys, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import oracle, oracle.sweeps as osw, oracle.blocksparse as ob
import pytenet_b200 as ptb
from pytenet_b200 import bond_ops, mps as pmps
warnings.simplefilter("ignore")
z = np.load(os.path.join(ROOT, "tests/golden/tdvp_xxz_L10.npz"))
n = int(z["h/nsites"])
w = [z[f"h/w{i}"] for i in range(n)]; wq = [z[f"h/qb{i}"] for i in range(n + 1)]
rec_o, rec_g = [], []
o_split = ob.split_block_sparse_matrix_svd
def o_probe(a, q0, q1, tol):
    u, s, v, q = ob.block_sparse_svd(a, q0, q1)
    rec_o.append((a.copy(), s.copy()))
    return o_split(a, q0, q1, tol)
osw.ob.split_block_sparse_matrix_svd = o_probe
g_split = bond_ops.split_block_sparse_matrix_svd
def g_probe(a, q0, q1, tol):
    u, s, v, q = ptb.block_sparse_svd(a, q0, q1)
    rec_err = (torch.linalg.norm((u * torch.as_tensor(s, device=u.device)) @ v - a) / torch.linalg.norm(a)).item()
    print(f"   device svd {tuple(a.shape)}: u conj={u.is_conj()} stride={u.stride()} | v conj={v.is_conj()} "
          f"stride={v.stride()} | recon err {rec_err:.1e}")
    rec_g.append((a.cpu().numpy().copy(), np.array(s)))
    return g_split(a, q0, q1, tol)
pmps.split_block_sparse_matrix_svd = g_probe
nst = int(sys.argv[1]) if len(sys.argv) > 1 else 1
psi_o = osw.Chain([z[f"psi0/a{i}"] for i in range(n)], z["psi0/qsite"], [z[f"psi0/qb{i}"] for i in range(n + 1)])
osw.tdvp_twosite(w, wq, psi_o, complex(z["dt"]), nst, numiter_lanczos=10, tol_split=1e-10)
h = ptb.MPO.from_tensors(z["h/qsite"], wq, w)
psi_g = ptb.MPS.from_tensors(z["psi0/qsite"], [z[f"psi0/qb{i}"] for i in range(n + 1)], [z[f"psi0/a{i}"] for i in range(n)])
ptb.tdvp_twosite(h, psi_g, complex(z["dt"]), nst, numiter_lanczos=10, tol_split=1e-10)
print("oracle dims", psi_o.bond_dims); print("device dims", psi_g.bond_dims)
for i, ((ao, so), (ag, sg)) in enumerate(zip(rec_o, rec_g)):
    same = ao.shape == ag.shape
    # gauge-invariant comparison of the matrices handed to the SVD: singular values
    m = min(len(so), len(sg))
    ds = np.max(np.abs(so[:m] - sg[:m]))
    po = (so / np.linalg.norm(so)) ** 2; pg = (sg / np.linalg.norm(sg)) ** 2
    # recompute singular values of the device matrix with LAPACK
    sl = np.linalg.svd(ag, compute_uv=False)
    print(f"split {i:2d} shape o{ao.shape} g{ag.shape} max|ds|={ds:.2e} lapack-vs-cusolver on device matrix: "
          f"{np.max(np.abs(sl[:len(sg)] - np.sort(sg)[::-1][:len(sl)])):.2e}  kept o/g: "
          f"{len(ob.retained_bond_indices(so,1e-10))}/{len(ob.retained_bond_indices(sg,1e-10))}")
    if not same or ds > 1e-9:
        print("   oracle s:", so[:20]); print("   device s:", sg[:20]); break
