"""C3 timing alone (dev helper): PopulationPredictiveModel.sample, 10k patients."""
import time, numpy as np, chi_b200
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
for k in range(6):
    t0 = time.perf_counter()
    out = ppm.sample(theta, times, n_samples=10000, seed=1, return_df=False)
    print('C3 call %d: %.1f ms' % (k, (time.perf_counter() - t0) * 1e3), np.isfinite(out).all())
