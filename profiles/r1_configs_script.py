import time, numpy as np, chi_b200
from chi_b200 import _synthetic
def timeit(f, n=3, warm=1):
    for _ in range(warm): f()
    t0=time.perf_counter()
    for _ in range(n): r=f()
    return (time.perf_counter()-t0)/n, r
# C3: erlotinib TGI, 10k virtual patients, PopulationPredictiveModel.sample
model = chi_b200.library.ModelLibrary().erlotinib_tumour_growth_inhibition_model()
model.set_administration('central', direct=True)
model.set_outputs(['global.tumour_volume'])
model.set_tolerance(abs_tol=1e-8, rel_tol=1e-6)
pred = chi_b200.PredictiveModel(model, [chi_b200.LogNormalErrorModel()])
pred.set_dosing_regimen(dose=5, start=0, duration=0.01, period=1, num=30)
pm = chi_b200.ComposedPopulationModel([chi_b200.PooledModel(1), chi_b200.LogNormalModel(1), chi_b200.PooledModel(2), chi_b200.LogNormalModel(3, centered=False), chi_b200.PooledModel(1)])
theta = [0.0, np.log(0.2), 0.2, 2.0, 1.0, np.log(0.5), np.log(0.3), np.log(0.1), 0.2, 0.3, 0.1, 0.1]
ppm = chi_b200.PopulationPredictiveModel(pred, pm)
times = np.arange(0.0, 31.0)
dt, out = timeit(lambda: ppm.sample(theta, times, n_samples=10000, seed=1, return_df=False))
print('C3: PopulationPredictiveModel.sample 10k patients x 31 times: %.1f ms (%.2e patients/s) shape %s finite %s' % (dt*1e3, 10000/dt, out.shape, np.isfinite(out).all()))
# C5: covariate population, CAM error, 100k individuals, evaluateS1
t0=time.perf_counter()
prob = _synthetic.two_compartment_problem(100000, seed=0, n_doses=3, error='cam', covariates=True, rtol=1e-6, atol=1e-8)
print('C5 setup (100k individuals, incl. synthetic data): %.1f s' % (time.perf_counter()-t0))
post = prob['log_posterior']
x = prob['x0']
dt, (s, g) = timeit(lambda: post.evaluateS1(x))
print('C5: HierarchicalLogPosterior.evaluateS1, 100k individuals, %d parameters: %.1f ms/eval (%.2e solves/s) score %.6g finite grad %s' % (len(x), dt*1e3, 100000/dt, s, np.isfinite(g).all()))
X = np.tile(x, (8,1))*(1+0.001*np.random.default_rng(0).normal(size=(8,len(x))))
dt, (s, g) = timeit(lambda: post.evaluateS1_batch(X, copy=False))
print('C5 batched x8: %.1f ms/step -> %.1f evals/s (%.2e solves/s)' % (dt*1e3, 8/dt, 8*100000/dt))
# C1: LogPosterior.evaluateS1 latency, 1 individual 100 times
m1 = chi_b200.library.ModelLibrary().one_compartment_pk_model(); m1.set_administration('central'); m1.set_dosing_regimen(10, 0, 0.01)
m1.set_tolerance(abs_tol=1e-8, rel_tol=1e-6)
t1 = np.linspace(0.1,10,100); y1 = np.abs(5*np.exp(-t1)+0.1*np.random.default_rng(1).normal(size=100))
ll = chi_b200.LogLikelihood(m1, chi_b200.GaussianErrorModel(), y1, t1)
lp = chi_b200.LogPosterior(ll, chi_b200.ComposedLogPrior(*[chi_b200.LogNormalLogPrior(0,1)]*4))
dt,_ = timeit(lambda: lp.evaluateS1([0.3,2.0,1.0,0.1]), n=50, warm=3)
print('C1: LogPosterior.evaluateS1 latency (1 individual, 100 times): %.3f ms' % (dt*1e3))
# C4: 1024 chains x 500 individuals, forward only (__call__ semantics of an MCMC iteration)
prob4 = _synthetic.two_compartment_problem(500, seed=0, n_doses=3, rtol=1e-6, atol=1e-8)
post4 = prob4['log_posterior']
X4 = np.tile(prob4['x0'], (1024, 1)) * (1 + 0.001 * np.random.default_rng(0).normal(size=(1024, len(prob4['x0']))))
dt, s4 = timeit(lambda: post4.evaluate_batch(X4), n=5, warm=2)
print('C4: HierarchicalLogPosterior.evaluate_batch, 1024 chains x 500 individuals: %.2f ms per MCMC iteration = %.0f evals/s = %.2e ODE solves/s, finite %s' % (dt*1e3, 1024/dt, 1024*500/dt, np.isfinite(np.asarray(s4)).all()))
