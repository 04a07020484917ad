import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from calibr8_b200 import meshgen, partition
from calibr8_b200.capi import Context
mesh = meshgen.box_tets(6, notch_radius=0.3)
_, part = partition.partition_mesh(mesh, 2, rank=0)
ctx = Context(0)
ctx.set_mesh(3, part.conn, part.coords)
ctx.set_model("mechanics", "small_J2", dict(E=1000., nu=.25, K=100., Y=2., cte=0., delta_T=0.), max_iters=50, abs_tol=1e-12, rel_tol=1e-12)
ctx.set_partition(part)
class Fake:
    def exchange(self, send, nb): return np.zeros((int(part.recv_ptr[-1]), nb))
    def allreduce(self, buf): return buf
ctx.set_comm_host(Fake())
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
x, xp = ctx.alloc("x"), ctx.alloc("x"); xi, xip = ctx.alloc("xi"), ctx.alloc("xi")
A, b = ctx.alloc("A"), ctx.alloc("b")
ctx.init_xi(xi); ctx.init_xi(xip)
x.copy_(torch.randn_like(x) * 1e-4)
print("nf", ctx.forward_jacobian(x, xp, xip, xi, A, b))
# identity on ghost-free diagonal for stability
rhs = torch.randn(ctx.n_dofs, dtype=torch.float64, device="cuda"); sol = torch.zeros_like(rhs)
print(ctx.gmres(A, rhs, sol, restart=50, max_iters=100, rel_tol=1e-8))
print(ctx.preconditioner_info())
torch.cuda.synchronize(); print("ok")
