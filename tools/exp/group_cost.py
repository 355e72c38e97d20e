"""experiment: does grouping the vectors of an individual by a cheap key reduce the synchronisation cost
(sum over stops of the max over the 32 lanes of a warp)?  host build"""
import ctypes, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import cport
import bench
n_ids, C, spread = 6, 1024, float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
times, events = bench.host_design(n_ids, 0, 10, 3)
key = cport.find_key('two_compartment_pk_model', 'central', True)
pop = cport.Population(key, ['central.drug_concentration'], ['lognormal'], times, [np.ones(10)] * n_ids, None, events)
rng = np.random.default_rng(1)
eta = rng.normal(size=(n_ids, 5)); mu = np.log(bench.TWO_CMT_TYPICAL)
X = np.zeros((C, n_ids, 8))
for l in range(C):
    m = mu + spread * rng.normal(size=5); sg = 0.3 * np.abs(1 + spread * rng.normal(size=5))
    e = eta + spread * rng.normal(size=(n_ids, 5))
    X[l, :, 2:7] = np.exp(m + sg * e); X[l, :, :2] = np.abs(0.1 * spread * rng.normal(size=2)); X[l, :, 7] = 0.1
x = X.reshape(-1, 8)
a = pop.args(); B = x.shape[0]
ll = np.zeros(B); grad = np.zeros((B, 8)); st = np.zeros(B, np.int32); ns = np.zeros(B, np.int32); nr = np.zeros(B, np.int32)
a.B = B; a.x = x.ctypes.data; a.x_stride = 8; a.mode = 1; a.with_sens = 1
a.rtol, a.atol, a.max_steps = 1e-8, 1e-10, 100000
a.ll = ll.ctypes.data; a.grad = grad.ctypes.data; a.grad_stride = 8
a.status = st.ctypes.data; a.n_steps = ns.ctypes.data; a.n_rej = nr.ctypes.data
a.refill_wait = 2**31 - 1
lib = ctypes.CDLL('/tmp/sync_cost.so'); MS = 40
out = np.zeros((B, MS), np.int32)
lib.trace(ctypes.byref(a), out.ctypes.data_as(ctypes.c_void_p), MS)
out = out.reshape(C, n_ids, MS)
def cost(order_per_id):
    tot = 0
    for i in range(n_ids):
        o = out[order_per_id[i], i, :].reshape(C // 32, 32, MS)
        tot += o.max(1).sum()
    return tot / (n_ids * C // 32)
mean = out.sum(-1).mean()
print('attempts mean %.2f' % mean)
print('random grouping      : iterations per warp %.2f  efficiency %.3f' % (cost([np.arange(C)] * n_ids), mean / cost([np.arange(C)] * n_ids)))
tot = out.sum(-1)
print('sorted by total count: %.2f  %.3f' % (cost([np.argsort(tot[:, i], kind='stable') for i in range(n_ids)]), mean / cost([np.argsort(tot[:, i], kind='stable') for i in range(n_ids)])))
psi = X[:, :, 2:7]   # size_c, ke, k12, k21, size_p
keys = {'ke + k12 + k21': psi[:, :, 1] + psi[:, :, 2] + psi[:, :, 3], 'ke': psi[:, :, 1], 'k12': psi[:, :, 2], 'k21': psi[:, :, 3]}
for name, k in keys.items():
    c = cost([np.argsort(k[:, i], kind='stable') for i in range(n_ids)])
    print('sorted by %-12s: %.2f  %.3f' % (name, c, mean / c))
