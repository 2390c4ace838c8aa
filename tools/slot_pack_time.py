"""Times the two slot_pack launches of the training step (first-block input, output-layer gradient) at 256 meshes."""
import sys, torch
sys.path.insert(0, '/root/repo')
from sdvae_b200 import cabi, fixtures as fx
from sdvae_b200.tables import spiral_table, restricted_spiral_table, pool_table
DEV='cuda:0'
tabs = fx.craniofacial_tables()
sp = [s.to(DEV) for s in tabs.spiral_tensors()]; dn=[d.to(DEV) for d in tabs.down_tensors()]
B=int(sys.argv[1]) if len(sys.argv) > 1 else 256; V=tabs.num_vertices[0]
full = spiral_table(sp[0]); sub = restricted_spiral_table(sp[0], pool_table(dn[0]))
x = torch.randn(B, V, 3, device=DEV)
P = torch.empty(B, sub.n_rows, 32, device=DEV); G = torch.empty(B, V, 32, device=DEV)
cp, cs = full.inverse()
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
print('slot_pack en0 (fwd table, %d rows): %.3f ms' % (sub.n_rows, t(lambda: cabi.slot_pack(x, None, sub.idx, P, B, V, sub.n_rows, 9, 3))))
print('slot_pack G (inverse table, %d rows): %.3f ms' % (V, t(lambda: cabi.slot_pack(x, cp, cs, G, B, V, V, 9, 3))))
