"""Dev helper (GPU): cProfile of one PopulationPredictiveModel.sample step of config C3."""
import cProfile, pstats, sys, time
import numpy as np
sys.path.insert(0, '.')
import chi_b200
model = chi_b200.library.ModelLibrary().erlotinib_tumour_growth_inhibition_model()
model.set_administration('central', direct=True)
model.set_outputs(['global.tumour_volume'])
model.set_tolerance(abs_tol=1e-10, rel_tol=1e-8)
pred = chi_b200.PredictiveModel(model, [chi_b200.LogNormalErrorModel()])
pred.set_dosing_regimen(dose=5, start=0, duration=0.01, period=1, num=30)
pm = chi_b200.ComposedPopulationModel([
    chi_b200.PooledModel(1), chi_b200.LogNormalModel(1), chi_b200.PooledModel(2),
    chi_b200.LogNormalModel(3, centered=False), chi_b200.PooledModel(1)])
theta = [0.0, np.log(0.2), 0.2, 2.0, 1.0, np.log(0.5), np.log(0.3), np.log(0.1), 0.2, 0.3, 0.1, 0.1]
ppm = chi_b200.PopulationPredictiveModel(pred, pm)
times = np.arange(0.0, 31.0)
step = lambda: ppm.sample(theta, times, n_samples=10000, seed=1, return_df=False)
for _ in range(3): step()
t0 = time.perf_counter()
for _ in range(5): step()
print('ms per step', (time.perf_counter() - t0) / 5 * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(5): step()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(25)
