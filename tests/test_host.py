"""Host-side logic and the C-ABI library without a GPU."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

import chi_b200
from chi_b200 import _build, _codegen, _lib, _library, _sbml
from oracle import cport, ode_exact

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = '/root/reference'


def test_library_exports_every_declared_symbol(native_lib):
    with open(os.path.join(ROOT, 'include', 'chi_b200.h')) as f:
        header = f.read()
    declared = set(re.findall(r'\b(chi_b200_[a-z0-9_]+)\s*\(', header))
    declared.discard('chi_b200_problem')
    assert declared == set(_lib.exported_symbols())
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), name
    assert native_lib.chi_b200_abi_version() == 1


def test_registry_holds_all_builtin_models(native_lib):
    with open(os.path.join(_build.GEN, 'index.json')) as f:
        index = json.load(f)
    assert native_lib.chi_b200_n_models() >= len(index) == 11
    for key, info in index.items():
        idx = native_lib.chi_b200_model_find(key.encode())
        assert idx >= 0
        out = np.zeros(8, dtype=np.int32)
        assert native_lib.chi_b200_model_info(
            idx, out.ctypes.data_as(_lib.c_i32p)) == 0
        assert out[0] == len(info['states'])
        assert out[1] == len(info['constants'])
        assert out[2] == len(info['outputs'])
        assert bool(out[3]) == info['linear']


def test_generated_sources_are_up_to_date():
    """The committed gen/*.cuh equal what the code generator emits."""
    for name, admin, model in _library.builtin_variants():
        code = _codegen.ModelCode(model)
        path = os.path.join(_build.GEN, 'model_%s.cuh' % code.key)
        with open(path) as f:
            assert f.read() == code.header(), (name, admin)


def test_no_cpu_fallback_without_gpu(native_lib):
    if native_lib.chi_b200_device_count() > 0:
        pytest.skip('a GPU is present')
    model = chi_b200.library.ModelLibrary().one_compartment_pk_model()
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        model.simulate([0, 2, 1], [1.0, 2.0])
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        chi_b200.GaussianErrorModel().compute_log_likelihood(
            [1.0], [1.0], [1.0])
    pm = chi_b200.LogNormalModel(2)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        pm.compute_log_likelihood([0.1, 0.2, 1, 1], np.ones((3, 2)))


def test_parameter_names_and_order():
    """chi/tests/test_mechanistic_models.py:127-142, 451-481."""
    lib = chi_b200.library.ModelLibrary()
    m = lib.tumour_growth_inhibition_model_koch()
    assert m.parameters() == [
        'global.tumour_volume', 'global.drug_concentration', 'global.kappa',
        'global.lambda_0', 'global.lambda_1']
    assert m.outputs() == ['global.tumour_volume']
    assert m.n_parameters() == 5 and m.n_outputs() == 1
    assert not m.supports_dosing()
    m = lib.one_compartment_pk_model()
    assert m.outputs() == ['central.drug_concentration']
    assert m.parameters() == [
        'central.drug_amount', 'central.size', 'global.elimination_rate']
    m.set_administration(compartment='central')
    assert m.parameters() == [
        'central.drug_amount', 'central.size', 'global.elimination_rate']
    m.set_administration(compartment='central', direct=False)
    assert m.parameters() == [
        'central.drug_amount', 'dose.drug_amount', 'central.size',
        'dose.absorption_rate', 'global.elimination_rate']
    assert m.n_parameters() == 5 and m.supports_dosing()
    m = lib.erlotinib_tumour_growth_inhibition_model()
    assert m.parameters() == [
        'central.drug_amount', 'global.tumour_volume', 'central.size',
        'global.critical_volume', 'global.elimination_rate', 'global.kappa',
        'global.lambda']


def test_administration_and_dosing_regimen_api():
    """chi/tests/test_mechanistic_models.py:483-593."""
    m = chi_b200.library.ModelLibrary().one_compartment_pk_model()
    with pytest.raises(ValueError, match='The route of administration'):
        m.set_dosing_regimen(dose=1)
    with pytest.raises(ValueError, match='The model does not'):
        m.set_administration(compartment='SOME COMP')
    with pytest.raises(ValueError, match='The drug amount variable'):
        m.set_administration(compartment='central', amount_var='SOME VAR')
    with pytest.raises(ValueError, match='The variable <'):
        m.set_administration(
            compartment='central', amount_var='drug_concentration')
    m.set_administration(compartment='central')
    m.set_dosing_regimen(1, 5)
    (e,) = m.dosing_regimen().events()
    assert (e.level(), e.start(), e.period(), e.duration(),
            e.multiplier()) == (100.0, 5, 0, 0.01, 0)
    m.set_dosing_regimen(1, 5, 1)
    (e,) = m.dosing_regimen().events()
    assert (e.level(), e.duration()) == (1.0, 1)
    m.set_dosing_regimen(1, 5, period=1)
    (e,) = m.dosing_regimen().events()
    assert (e.period(), e.multiplier()) == (1, 0)
    m.set_dosing_regimen(10, 3, 0.01, 5, 4)
    (e,) = m.dosing_regimen().events()
    assert (e.level(), e.start(), e.period(), e.multiplier()) == (
        1000.0, 3, 5, 4)
    regimen = chi_b200.blocktrain(period=1, duration=0.1, limit=1)
    m.set_dosing_regimen(regimen)
    (e,) = m.dosing_regimen().events()
    assert (e.level(), e.start(), e.period(), e.duration(),
            e.multiplier()) == (1, 0, 1, 0.1, 1)
    assert m.administration() == {'compartment': 'central', 'direct': True}


def test_outputs_sensitivities_and_copy_semantics():
    """Setting outputs / administration disables sensitivities
    (chi/tests/test_mechanistic_models.py:405-438); copy resets them."""
    m = chi_b200.library.ModelLibrary().one_compartment_pk_model()
    m.enable_sensitivities(True)
    assert m.has_sensitivities()
    m.set_outputs(['central.drug_amount', 'central.drug_concentration'])
    assert not m.has_sensitivities() and m.n_outputs() == 2
    with pytest.raises(KeyError, match='does not exist'):
        m.set_outputs(['nope.nope'])
    with pytest.raises(ValueError, match='None of the parameters'):
        m.enable_sensitivities(True, ['nope'])
    m.enable_sensitivities(True, ['central.size'])
    c = m.copy()
    assert not c.has_sensitivities() and m.has_sensitivities()
    m.set_output_names({'central.drug_amount': 'A'})
    assert m.outputs()[0] == 'A'
    m.set_parameter_names({'central.size': 'V'})
    assert m.parameters()[1] == 'V'
    with pytest.raises(TypeError):
        m.set_parameter_names(['V'])


def test_reduced_mechanistic_model_bookkeeping():
    m = chi_b200.library.ModelLibrary().one_compartment_pk_model()
    r = chi_b200.ReducedMechanisticModel(m)
    r.fix_parameters({'central.drug_amount': 0.0})
    assert r.n_parameters() == 2 and r.n_fixed_parameters() == 1
    assert r.parameters() == ['central.size', 'global.elimination_rate']
    r.fix_parameters({'central.drug_amount': None})
    assert r.n_parameters() == 3
    with pytest.raises(TypeError):
        chi_b200.ReducedMechanisticModel('model')


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason='no reference')
def test_sbml_reader_reproduces_builtin_models():
    files = {
        'one_compartment_pk_model':
            'chi/library/model_library/pk_one_comp.xml',
        'two_compartment_pk_model':
            'docs/source/getting_started/code/two_compartment_pk_model.xml',
        'erlotinib_tumour_growth_inhibition_model':
            'chi/library/model_library/temporary_full_pkpd_model.xml',
        'tumour_growth_inhibition_model_koch':
            'chi/library/model_library/tgi_Koch_2009.xml',
        'tumour_growth_inhibition_model_koch_reparametrised':
            'chi/library/model_library/tgi_Koch_2009_reparametrised.xml'}
    for name, admin, model in _library.builtin_variants():
        s = _sbml.read_sbml(os.path.join(REFERENCE, files[name]))
        if admin:
            state = admin[0] + '.drug_amount'
            if not admin[1]:
                state = s.add_dose_compartment(state)
            s.add_dose_rate(state)
        assert _codegen.ModelCode(s).key == _codegen.ModelCode(model).key
    # the docs' compartment-free model: everything lives in `global`
    s = _sbml.read_sbml(os.path.join(
        REFERENCE,
        'docs/source/getting_started/code/one_compartment_pk_model.xml'))
    assert s.states == ['global.drug_amount']
    assert sorted(s.constants) == [
        'global.elimination_rate', 'global.volume']
    assert list(s.intermediates) == ['global.drug_concentration']


def test_population_model_bookkeeping():
    """SURVEY.md section 8a worked layouts."""
    pm = chi_b200.ComposedPopulationModel([
        chi_b200.PooledModel(2), chi_b200.LogNormalModel(5, centered=False),
        chi_b200.PooledModel(1)])
    pm.set_n_ids(3)
    assert pm.n_hierarchical_parameters(3) == (15, 13)
    assert pm.get_special_dims() == (
        [[0, 2, 0, 2, True], [7, 8, 12, 13, True]], 3, 0)
    cpm = chi_b200.ComposedPopulationModel([
        chi_b200.PooledModel(2),
        chi_b200.CovariatePopulationModel(
            chi_b200.LogNormalModel(5, centered=False),
            chi_b200.LinearCovariateModel(n_cov=2)),
        chi_b200.PooledModel(2)])
    cpm.set_n_ids(3)
    assert cpm.n_hierarchical_parameters(3) == (15, 34)
    assert cpm.n_covariates() == 2
    assert len(cpm.get_parameter_names()) == 34
    het = chi_b200.ComposedPopulationModel([
        chi_b200.HeterogeneousModel(1), chi_b200.GaussianModel(1)])
    het.set_n_ids(4)
    assert het.n_hierarchical_parameters(4) == (4, 6)
    assert het.get_special_dims() == ([[0, 1, 0, 4, False]], 0, 1)
    with pytest.raises(TypeError):
        chi_b200.ComposedPopulationModel(['x'])
    with pytest.raises(ValueError):
        chi_b200.LogNormalModel(0)


def test_population_sampling_matches_numpy_stream():
    """sample() draws on the host in the reference's order
    (chi/_population_models.py:1797-1846, 2678-2732, 793-872)."""
    pm = chi_b200.ComposedPopulationModel([
        chi_b200.PooledModel(1), chi_b200.LogNormalModel(2),
        chi_b200.GaussianModel(1, centered=False)])
    theta = [3.0, 0.1, 0.2, 0.3, 0.4, 1.0, 2.0]
    s = pm.sample(theta, n_samples=4, seed=7)
    rng = np.random.default_rng(7)
    ref_ln = rng.lognormal(mean=[0.1, 0.2], sigma=[0.3, 0.4], size=(4, 2))
    ref_g = rng.normal(loc=[0.0], scale=[1.0], size=(4, 1))
    np.testing.assert_array_equal(s[:, 0], 3.0)
    np.testing.assert_array_equal(s[:, 1:3], ref_ln)
    np.testing.assert_array_equal(s[:, 3:], ref_g)


def test_error_model_sampling_matches_numpy_stream():
    """chi/tests/test_error_models.py:284-315 style: seed-pinned draws."""
    y = np.array([1.0, 2.0, 3.0])
    s = chi_b200.GaussianErrorModel().sample([0.5], y, n_samples=2, seed=1)
    ref = y[:, None] + np.random.default_rng(1).normal(0, 0.5, size=(3, 2))
    np.testing.assert_array_equal(s, ref)
    s = chi_b200.ConstantAndMultiplicativeGaussianErrorModel().sample(
        [0.1, 0.2], y, n_samples=2, seed=3)
    rng = np.random.default_rng(3)
    b = rng.normal(0, 0.1, size=(3, 2))
    r = rng.normal(0, 0.2, size=(3, 2))
    np.testing.assert_array_equal(s, y[:, None] + b + y[:, None] * r)
    with pytest.raises(ValueError, match='number of provided parameters'):
        chi_b200.GaussianErrorModel().sample([1, 2], y)


def test_log_likelihood_input_validation():
    """chi/tests/test_log_pdfs.py:1720-1768 message prefixes."""
    m = chi_b200.library.ModelLibrary().one_compartment_pk_model()
    em = chi_b200.GaussianErrorModel()
    with pytest.raises(TypeError, match='The mechanistic model'):
        chi_b200.LogLikelihood('m', em, [1, 2], [1, 2])
    with pytest.raises(ValueError, match='One error model has to be'):
        chi_b200.LogLikelihood(m, [em, em], [1, 2], [1, 2])
    with pytest.raises(TypeError, match='The error models have to'):
        chi_b200.LogLikelihood(m, ['e'], [1, 2], [1, 2])
    with pytest.raises(ValueError, match='Times cannot be negative'):
        chi_b200.LogLikelihood(m, em, [1, 2], [-1, 2])
    with pytest.raises(ValueError, match='Times must be increasing'):
        chi_b200.LogLikelihood(m, em, [1, 2], [2, 1])
    with pytest.raises(ValueError, match='The observations and times'):
        chi_b200.LogLikelihood(m, em, [1, 2, 3], [1, 2])
    ll = chi_b200.LogLikelihood(m, em, [1, 2], [1, 2])
    assert ll.n_parameters() == 4 and ll.n_observations() == [2]
    assert ll.get_parameter_names()[-1] == 'Sigma'


def test_cpu_port_converges_to_exact_solution():
    """The host build of the kernel's per-system routine tracks the exact
    solution as the tolerance tightens (step control sanity, no GPU)."""
    key = cport.find_key('two_compartment_pk_model', 'central', True)
    times = np.sort(np.random.default_rng(1).uniform(0.25, 24, 10))
    reg = [(10.0, 0.0, 1.0, 8.0, 3)]
    psi = [0, 0, 2, 1, 5, 1, 10]
    pop = cport.Population(
        key, ['central.drug_concentration'], None, [times], None, None, [reg])
    m = ode_exact.two_compartment(True)
    ye, se = ode_exact.simulate(m, psi, times, reg)
    prev = None
    for rtol in (1e-6, 1e-8, 1e-10):
        y, s, status, nst, nrj = pop.simulate(psi, rtol=rtol, atol=rtol * 0.01)
        assert status[0] == 0
        err = np.max(np.abs(y - ye[0]) / np.max(np.abs(ye[0])))
        scale = np.max(np.abs(se[:, 0, :]), 0)
        scale[scale == 0] = 1.0   # d/d(central.size) of an amount is zero
        serr = np.max(np.abs(s - se[:, 0, :]) / scale)
        assert err < rtol and serr < 2 * rtol
        assert prev is None or nst[0] > prev
        prev = nst[0]


def test_cpu_port_rodas5_needs_fewer_steps_and_converges():
    """RODAS5 (order 5(4)) in the host build: converges to the exact solution
    and needs about half the steps of RODAS4 at the same tolerance."""
    key = cport.find_key('two_compartment_pk_model', 'central', True)
    times = np.sort(np.random.default_rng(1).uniform(0.25, 24, 10))
    reg = [(10.0, 0.0, 1.0, 8.0, 3)]
    psi = [0, 0, 2, 1, 5, 1, 10]
    pop = cport.Population(
        key, ['central.drug_concentration'], None, [times], None, None, [reg])
    m = ode_exact.two_compartment(True)
    ye, se = ode_exact.simulate(m, psi, times, reg)
    scale = np.max(np.abs(se[:, 0, :]), 0)
    scale[scale == 0] = 1.0
    errs = {}
    for rtol in (1e-6, 1e-8, 1e-10):
        y4, s4, st4, n4, _ = pop.simulate(
            psi, rtol=rtol, atol=rtol * 0.01, stream_cb=1)
        y5, s5, st5, n5, _ = pop.simulate(
            psi, rtol=rtol, atol=rtol * 0.01, stream_cb=5)
        assert st4[0] == 0 and st5[0] == 0
        assert n5[0] < 0.65 * n4[0]
        err = np.max(np.abs(y5 - ye[0]) / np.max(np.abs(ye[0])))
        serr = np.max(np.abs(s5 - se[:, 0, :]) / scale)
        assert err < rtol and serr < 2 * rtol
        errs[rtol] = err
    # two decades of tolerance buy about two decades of accuracy
    assert errs[1e-10] < 1e-3 * errs[1e-6]
    # the streamed and the lane formulation of RODAS4 are the same arithmetic
    y0, s0, _, n0, _ = pop.simulate(psi, rtol=1e-8, atol=1e-10, stream_cb=0)
    y1, s1, _, n1, _ = pop.simulate(psi, rtol=1e-8, atol=1e-10, stream_cb=1)
    assert n0[0] == n1[0]
    np.testing.assert_allclose(y0, y1, rtol=1e-12)
    np.testing.assert_allclose(s0, s1, rtol=1e-9, atol=1e-13)


def test_bench_host_design_matches_synthetic():
    import bench
    from chi_b200 import _synthetic
    t1, e1 = bench.host_design(7, 3, 10, 3)
    t2, r2 = _synthetic.two_compartment_design(7, 3, 10, 3)
    for a, b in zip(t1, t2):
        np.testing.assert_array_equal(a, b)
    for a, b in zip(e1, r2):
        np.testing.assert_array_equal(np.array(a), b.as_array())
    np.testing.assert_array_equal(
        bench.TWO_CMT_TYPICAL, _synthetic.TWO_CMT_TYPICAL)
    assert (bench.BENCH_RTOL, bench.BENCH_ATOL) == (
        _synthetic.BENCH_RTOL, _synthetic.BENCH_ATOL)


def test_k_stage_tables_follow_from_the_transformed_tables():
    """The K_ij constants of the linear-model bodies (k-stage form,
    chi_b200/csrc/integrator_stream.cuh) are (alpha + Gamma_offdiag) / gamma
    of the same method whose transformed tables (a_ij, C_ij, gamma) the
    nonlinear bodies use: re-derive them (tools/derive_kform.py) and compare;
    also the structure the RODAS5 body relies on -- K72 = K82 = 0, row 8 = row
    7 + two entries -- and that one k-form step of a random linear system
    reproduces the transformed form."""
    import importlib.util
    import mpmath as mp
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location(
        'derive_kform', os.path.join(root, 'tools', 'derive_kform.py'))
    dk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(dk)
    for name, path, s in (
            ('rodas4', 'chi_b200/csrc/integrator.cuh', 6),
            ('rodas5', 'chi_b200/csrc/integrator_stream.cuh', 8)):
        vals = dk.read_namespace(os.path.join(root, path), name)
        g, alpha, G, beta = dk.derive(vals, s)
        K = np.array([[float(beta[i, j] / g) for j in range(s)]
                      for i in range(s)])
        typed = np.zeros((s, s))
        for i in range(1, s):
            for j in range(i):
                key = 'K%d%d' % (i + 1, j + 1)
                if key in vals:
                    typed[i, j] = float(vals[key])
        if name == 'rodas5':
            assert abs(K[6, 1]) < 1e-13 and abs(K[7, 1]) < 1e-13
            np.testing.assert_allclose(K[7, :5], K[6, :5], atol=1e-13)
            typed[7, :5] = typed[6, :5]
            typed[7, 5] = typed[6, 5] + float(vals['K8D6'])
            K[6, 1] = K[7, 1] = 0.0
        np.testing.assert_allclose(typed, K, rtol=1e-13, atol=1e-14)
        # one step of z' = J z + f0, transformed form vs k-stage form
        a = np.zeros((s, s))
        C = np.zeros((s, s))
        for i in range(1, s):
            have = ('A%d1' % (i + 1)) in vals
            for j in range(i):
                C[i, j] = float(vals['C%d%d' % (i + 1, j + 1)])
                a[i, j] = float(vals['A%d%d' % (i + 1, j + 1)]) if have \
                    else (a[i - 1, j] if j < i - 1 else 1.0)
        rng = np.random.default_rng(0)
        n, h, gam = 3, 0.37, float(g)
        J = rng.normal(size=(n, n))
        f0 = rng.normal(size=n)
        y = rng.normal(size=n)
        Wi = np.linalg.inv(np.eye(n) / (gam * h) - J)
        U = np.zeros((s, n))
        for i in range(s):
            Z = y + a[i, :i] @ U[:i]
            U[i] = Wi @ (J @ Z + f0 + (C[i, :i] @ U[:i]) / h)
        y_new, err = Z + U[s - 1], U[s - 1]
        E = Wi @ J
        kap = np.zeros((s, n))
        V = np.zeros((s, n))
        for i in range(s):
            arg = y + typed[i, :i] @ kap[:i]
            kap[i] = E @ arg + Wi @ f0
            V[i] = arg + kap[i]
        np.testing.assert_allclose(V[s - 1], y_new, rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(V[s - 1] - V[s - 2], err, atol=1e-12)


def test_flop_model_of_the_bench_counts_the_k_stage_form():
    """bench.flops_per_step for the two-compartment model (DESIGN.md,
    'Algorithmic flops'): N = 2, 5 integrated columns of which 2 are
    initial-value columns; info = chi_b200_model_info."""
    import bench
    info = np.array([2, 5, 4, 3, 6, 3, 10, 0])   # linear, rank-one terms
    assert bench.flops_per_step(2, 5, info, stages=8, n_init=2) == 1048
    # ROSL(11), the default method of linear models: 21 (T, E) + 140
    # (state) + 2 x 112 (initial-value columns) + 3 x 162 (constants)
    assert bench.flops_per_step(
        2, 5, info, stages=11, n_init=2, method='ROSL11') == 871
    assert bench.flops_per_step(
        2, 0, info, stages=11, n_init=0, method='ROSL11') == 141
    assert bench.flops_per_step(2, 0, info, stages=8, n_init=2) == 206
    dense = info.copy()
    dense[3] = 1                                  # linear, dense column terms
    assert bench.flops_per_step(2, 5, dense, stages=8, n_init=2) == 1188


TRANSIT_SBML = """<?xml version="1.0" encoding="UTF-8"?>
<sbml xmlns="http://www.sbml.org/sbml/level3/version2/core" level="3" version="2">
  <model id="transit_pk" timeUnits="day">
    <listOfCompartments>
      <compartment id="central" name="central" size="1"/>
      <compartment id="gut" name="gut" size="1"/>
      <compartment id="transit" name="transit" size="1"/>
    </listOfCompartments>
    <listOfSpecies>
      <species id="drug" name="drug" compartment="central" initialAmount="0" hasOnlySubstanceUnits="true"/>
      <species id="a1" name="a1" compartment="gut" initialAmount="0" hasOnlySubstanceUnits="true"/>
      <species id="a2" name="a2" compartment="transit" initialAmount="0" hasOnlySubstanceUnits="true"/>
    </listOfSpecies>
    <listOfParameters>
      <parameter id="k_tr" value="1" constant="true"/>
      <parameter id="k_e" value="1" constant="true"/>
    </listOfParameters>
    <listOfReactions>
      <reaction id="r1" reversible="false">
        <listOfReactants><speciesReference species="a1"/></listOfReactants>
        <listOfProducts><speciesReference species="a2"/></listOfProducts>
        <kineticLaw><math xmlns="http://www.w3.org/1998/Math/MathML">
          <apply><times/><ci> k_tr </ci><ci> a1 </ci></apply></math></kineticLaw>
      </reaction>
      <reaction id="r2" reversible="false">
        <listOfReactants><speciesReference species="a2"/></listOfReactants>
        <listOfProducts><speciesReference species="drug"/></listOfProducts>
        <kineticLaw><math xmlns="http://www.w3.org/1998/Math/MathML">
          <apply><times/><ci> k_tr </ci><ci> a2 </ci></apply></math></kineticLaw>
      </reaction>
      <reaction id="r3" reversible="false">
        <listOfReactants><speciesReference species="drug"/></listOfReactants>
        <kineticLaw><math xmlns="http://www.w3.org/1998/Math/MathML">
          <apply><times/><ci> k_e </ci><ci> drug </ci></apply></math></kineticLaw>
      </reaction>
    </listOfReactions>
  </model>
</sbml>
"""


def test_cpu_port_linear_model_with_dense_column_terms(tmp_path):
    """A linear model in which one constant multiplies two states (transit
    compartments sharing a rate constant): its column term D_k has two
    non-zero columns, so it takes the k-stage form with the dense
    (W^-1 D_k) V product -- none of the built-in models does.  SBML -> code
    generator -> host build of the kernel's routine -> exact solution."""
    from chi_b200 import _build
    from oracle.csolver import build as port_build
    path = tmp_path / 'transit.xml'
    path.write_text(TRANSIT_SBML)
    model = chi_b200.PKPDModel(str(path))
    model.set_administration('gut', amount_var='a1_amount', direct=True)
    assert model.parameters() == [
        'central.drug_amount', 'gut.a1_amount', 'transit.a2_amount',
        'central.size', 'global.k_e', 'global.k_tr', 'gut.size',
        'transit.size']
    code = model.model_code()
    assert code.linear
    header = code.header()
    assert 'COL_RANK1 = false' in header
    gen = tmp_path / 'gen'
    gen.mkdir()
    paths = _build.write_model_sources(code, str(gen))
    hdr = [p for p in paths if p.endswith('.cuh')][0]
    port = cport.open_port(port_build.build_models({code.key: hdr},
                                                   str(tmp_path)))
    info = {'states': code.state_names, 'constants': code.const_names,
            'outputs': code.output_names}
    times = np.array([0.2, 0.5, 1.0, 1.7, 2.5, 4.0, 6.5, 9.0])
    reg = [(6.0, 0.0, 0.5, 3.0, 3)]
    psi = np.array([0.3, 0.1, 0.2, 2.0, 0.7, 1.9, 1.0, 1.0])
    pop = cport.Population(code.key, ['central.drug_amount'], None, [times],
                           None, None, [reg], info=info, port=port)
    omodel = ode_exact.OdeModel(
        ['central.drug_amount', 'gut.a1_amount', 'transit.a2_amount'],
        ['central.size', 'global.k_e', 'global.k_tr', 'gut.size',
         'transit.size'],
        {'central.drug_amount':
            'global__k_tr * transit__a2_amount'
            ' - global__k_e * central__drug_amount',
         'gut.a1_amount': '-global__k_tr * gut__a1_amount',
         'transit.a2_amount':
            'global__k_tr * gut__a1_amount'
            ' - global__k_tr * transit__a2_amount'},
        {'central.drug_amount': 'central__drug_amount'}, 'gut.a1_amount')
    omodel.set_outputs(['central.drug_amount'])
    assert omodel.parameter_names == model.parameters()
    y_ref, s_ref = ode_exact.simulate(omodel, psi, times, reg)
    scale = np.max(np.abs(s_ref[:, 0, :]), 0)
    scale[scale == 0] = 1.0
    for method in (11, 9, 5, 1):
        y, s, status, nst, _ = pop.simulate(
            psi, rtol=1e-10, atol=1e-12, stream_cb=method)
        assert status[0] == 0
        assert np.max(np.abs(y - y_ref[0])) / np.max(np.abs(y_ref[0])) < 1e-9
        assert np.max(np.abs(s - s_ref[:, 0, :]) / scale) < 1e-8


def test_cpu_port_stiff_parameter_sweep():
    """The k-stage form over six decades of rate constants (stiffness ratios
    up to 1e6 with steps far outside the explicit stability range): no
    integrator failure, and trajectories / sensitivities within the mixed
    tolerance of the exact solution."""
    key = cport.find_key('two_compartment_pk_model', 'central', True)
    times = np.sort(np.random.default_rng(1).uniform(0.25, 24, 10))
    reg = [(10.0, 0.0, 1.0, 8.0, 3)]
    pop = cport.Population(
        key, ['central.drug_concentration'], None, [times], None, None, [reg])
    m = ode_exact.two_compartment(True)
    rng = np.random.default_rng(7)
    rtol, atol = 1e-6, 1e-8
    for _ in range(120):
        psi = np.array([
            rng.uniform(0, 5), rng.uniform(0, 5), 10 ** rng.uniform(-1, 2),
            10 ** rng.uniform(-3, 3), 10 ** rng.uniform(-3, 3),
            10 ** rng.uniform(-3, 3), 10 ** rng.uniform(-1, 2)])
        for method in (11, 9, 5, 1):
            y, s, status, nst, nrj = pop.simulate(
                psi, rtol=rtol, atol=atol, stream_cb=method)
            assert status[0] == 0, (psi, method)
            assert nst[0] + nrj[0] < 2000
            ye, se = ode_exact.simulate(m, psi, times, reg)
            tol_y = 20 * (atol + rtol * np.max(np.abs(ye[0])))
            assert np.max(np.abs(y - ye[0])) < tol_y, (psi, method)
            tol_s = 20 * (atol + rtol * np.max(np.abs(se[:, 0, :]), 0))
            assert np.all(np.abs(s - se[:, 0, :]) < tol_s[None, :]), (
                psi, method)


CLEARANCE_SBML = """<?xml version="1.0" encoding="UTF-8"?>
<sbml xmlns="http://www.sbml.org/sbml/level3/version2/core" level="3" version="2">
  <model id="clearance_pk" timeUnits="day">
    <listOfCompartments>
      <compartment id="central" name="central" size="1"/>
    </listOfCompartments>
    <listOfSpecies>
      <species id="drug" name="drug" compartment="central" initialAmount="0" hasOnlySubstanceUnits="true"/>
    </listOfSpecies>
    <listOfParameters>
      <parameter id="clearance" value="1" constant="true"/>
      <parameter id="v_d" value="1" constant="true"/>
    </listOfParameters>
    <listOfReactions>
      <reaction id="elimination" reversible="false">
        <listOfReactants><speciesReference species="drug"/></listOfReactants>
        <kineticLaw><math xmlns="http://www.w3.org/1998/Math/MathML">
          <apply><divide/>
            <apply><times/><ci> clearance </ci><ci> drug </ci></apply>
            <ci> v_d </ci>
          </apply></math></kineticLaw>
      </reaction>
    </listOfReactions>
  </model>
</sbml>
"""


def test_cpu_port_rank_one_terms_that_depend_on_constants(tmp_path):
    """Clearance parametrisation dA/dt = -(CL / V_d) A + u: both column terms
    are rank one with a coefficient that is an expression of the constants
    (d_CL = -1/V_d, d_Vd = CL/V_d^2), not a literal as in the mass-action
    library models."""
    from chi_b200 import _build
    from oracle.csolver import build as port_build
    path = tmp_path / 'clearance.xml'
    path.write_text(CLEARANCE_SBML)
    model = chi_b200.PKPDModel(str(path))
    model.set_administration('central', direct=True)
    assert model.parameters() == [
        'central.drug_amount', 'central.size', 'global.clearance',
        'global.v_d']
    code = model.model_code()
    header = code.header()
    assert code.linear and 'COL_RANK1 = true' in header
    gen = tmp_path / 'gen'
    gen.mkdir()
    paths = _build.write_model_sources(code, str(gen))
    hdr = [p for p in paths if p.endswith('.cuh')][0]
    port = cport.open_port(port_build.build_models({code.key: hdr},
                                                   str(tmp_path)))
    info = {'states': code.state_names, 'constants': code.const_names,
            'outputs': code.output_names}
    times = np.array([0.1, 0.4, 1.0, 1.6, 2.5, 4.0, 7.0])
    reg = [(5.0, 0.0, 0.25, 2.0, 3)]
    psi = np.array([0.4, 1.0, 1.3, 2.2])
    pop = cport.Population(code.key, ['central.drug_amount'], None, [times],
                           None, None, [reg], info=info, port=port)
    omodel = ode_exact.OdeModel(
        ['central.drug_amount'],
        ['central.size', 'global.clearance', 'global.v_d'],
        {'central.drug_amount':
            '-global__clearance / global__v_d * central__drug_amount'},
        {'central.drug_amount': 'central__drug_amount'},
        'central.drug_amount')
    omodel.set_outputs(['central.drug_amount'])
    assert omodel.parameter_names == model.parameters()
    y_ref, s_ref = ode_exact.simulate(omodel, psi, times, reg)
    scale = np.max(np.abs(s_ref[:, 0, :]), 0)
    scale[scale == 0] = 1.0
    for method in (11, 5):
        y, s, status, _, _ = pop.simulate(psi, rtol=1e-10, atol=1e-12,
                                          stream_cb=method)
        assert status[0] == 0
        assert np.max(np.abs(y - y_ref[0])) / np.max(np.abs(y_ref[0])) < 1e-9
        assert np.max(np.abs(s - s_ref[:, 0, :]) / scale) < 1e-8


def test_batched_posterior_combines_prior_and_likelihood_on_the_host():
    """HierarchicalLogPosterior.evaluate_batch / evaluateS1_batch: the host
    prior runs on a worker thread while the likelihood call is in flight;
    rows whose prior is -inf get score -inf and gradient +inf
    (chi/_log_pdfs.py:463-498), with a batched prior and with a prior that
    only offers the single-vector calls.  The GPU likelihood is replaced by a
    stub here: this pins the host logic only."""
    from chi_b200 import _log_pdfs

    n_bottom, n_top, C = 6, 3, 5

    class StubLikelihood(object):
        def evaluate_batch(self, X, copy=True):
            return np.sum(X, axis=1)

        def evaluateS1_batch(self, X, copy=True):
            return np.sum(X, axis=1), np.ones_like(X) * 2.0

    class SinglePrior(object):
        """log p = -0.5 |t|^2, -inf when t[0] < 0 (no batch methods)."""
        def n_parameters(self):
            return n_top

        def __call__(self, t):
            return -np.inf if t[0] < 0 else -0.5 * float(np.dot(t, t))

        def evaluateS1(self, t):
            if t[0] < 0:
                return -np.inf, np.full(n_top, np.nan)
            return -0.5 * float(np.dot(t, t)), -np.asarray(t, float)

    class BatchPrior(SinglePrior):
        def evaluateS1_batch(self, T):
            p = np.array([self(t) for t in T])
            s = np.array([self.evaluateS1(t)[1] for t in T])
            return p, s

    rng = np.random.default_rng(3)
    X = rng.uniform(0.1, 1.0, size=(C, n_bottom + n_top))
    X[2, n_bottom] = -1.0
    for prior in (SinglePrior(), BatchPrior()):
        post = object.__new__(_log_pdfs.HierarchicalLogPosterior)
        post._log_likelihood = StubLikelihood()
        post._log_prior = prior
        post._n_bottom = n_bottom
        post._n_parameters = n_bottom + n_top
        score = post.evaluate_batch(X)
        s1, grad = post.evaluateS1_batch(X)
        expect = np.sum(X, axis=1) - 0.5 * np.sum(X[:, n_bottom:] ** 2, 1)
        ok = np.arange(C) != 2
        np.testing.assert_allclose(score[ok], expect[ok], rtol=1e-14)
        np.testing.assert_allclose(s1[ok], expect[ok], rtol=1e-14)
        assert np.isneginf(score[2]) and np.isneginf(s1[2])
        assert np.all(np.isposinf(grad[2]))
        np.testing.assert_allclose(grad[ok, :n_bottom], 2.0)
        np.testing.assert_allclose(
            grad[ok, n_bottom:], 2.0 - X[ok, n_bottom:], rtol=1e-14)
        assert np.all(np.isfinite(grad[ok]))


def test_invalid_prior_never_touches_other_shards_columns():
    """Individuals sharded over ranks: the likelihood's page-locked gradient
    buffer is reused across evaluations and only the rank's own columns are
    rewritten.  A -inf prior fills the gradient row with +inf -- in the owned
    columns only; the next evaluation must find the other individuals'
    columns still zero (stale inf entries were all-reduced)."""
    from chi_b200 import _log_pdfs

    n_ids, n_h, n_top, C = 6, 2, 3, 4
    n_bottom = n_ids * n_h
    shard = (2, 4)

    class ShardedStub(object):
        """Mimics chi_b200_hier_logpdf with a shard: one cached buffer, only
        the shard's bottom block and the top level are written."""
        def __init__(self):
            self.grad = np.zeros((C, n_bottom + n_top))
            self.score = np.zeros(C)

        def evaluateS1_batch(self, X, copy=True):
            self.score[:] = np.sum(X, axis=1)
            self.grad[:, shard[0] * n_h:shard[1] * n_h] = 2.0
            self.grad[:, n_bottom:] = 3.0
            return self.score, self.grad

        def _fill_rows(self, grad, rows, value):
            _log_pdfs.fill_owned_columns(
                grad, rows, value, shard, n_ids, n_bottom)

    class Prior(object):
        def n_parameters(self):
            return n_top

        def evaluateS1_batch(self, T):
            p = np.where(T[:, 0] < 0, -np.inf, -0.5 * np.sum(T * T, axis=1))
            return p, -T

    post = object.__new__(_log_pdfs.HierarchicalLogPosterior)
    post._log_likelihood = ShardedStub()
    post._log_prior = Prior()
    post._n_bottom = n_bottom
    post._n_parameters = n_bottom + n_top
    X = np.random.default_rng(0).uniform(0.1, 1.0, (C, n_bottom + n_top))
    X[1, n_bottom] = -1.0
    s, g = post.evaluateS1_batch(X, copy=False)
    assert np.isneginf(s[1])
    assert np.all(np.isposinf(g[1, 4:8])) and np.all(np.isposinf(g[1, 12:]))
    assert np.all(g[1, :4] == 0.0) and np.all(g[1, 8:12] == 0.0)
    X[1, n_bottom] = 0.5
    s, g = post.evaluateS1_batch(X, copy=False)
    assert np.all(np.isfinite(s)) and np.all(np.isfinite(g))
    assert np.all(g[:, :4] == 0.0) and np.all(g[:, 8:12] == 0.0)
    np.testing.assert_allclose(g[:, 4:8], 2.0)


def test_rosl_tables_order_and_stability():
    """ROSL(9) / ROSL(11) (integrator_stream.cuh, namespace rosl): the tables
    in the header are the ones tools/derive_rosl.py derives; stability
    function of order s-1, embedded method (last two terms dropped) of order
    s-2, |R(iy)| <= 1 on the imaginary axis, R(-inf) = 0."""
    import re
    sys_path = os.path.join(ROOT, 'tools')
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        'derive_rosl', os.path.join(sys_path, 'derive_rosl.py'))
    derive = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(derive)
    text = open(os.path.join(
        ROOT, 'chi_b200', 'csrc', 'integrator_stream.cuh')).read()
    for s_, gamma in ((9, '0.2'), (11, '0.16')):
        c, order, order_hat, amax, rmax, rinf = derive.check(
            s_, derive.mp.mpf(gamma))
        assert abs(order - (s_ - 1)) < 0.1 and abs(order_hat - (s_ - 2)) < 0.1
        assert amax <= 1 + 1e-12 and rmax <= 1 + 1e-12 and rinf < 1e-6
        block = text[text.index('struct Table<%d>' % s_):]
        block = block[block.index('C[%d] = {' % s_):block.index('};')]
        vals = [float(v) for v in re.findall(
            r'-?\d+\.\d+(?:[eE][-+]?\d+)?', block.split('{', 1)[1])]
        assert len(vals) == s_
        np.testing.assert_allclose(vals, [float(v) for v in c], rtol=1e-15)
        # the device copy in constant memory carries the same numbers
        dev = text[text.index('static __constant__ double C%d[%d]' % (s_, s_)):]
        dev = dev[:dev.index('};')]
        dvals = [float(v) for v in re.findall(
            r'-?\d+\.\d+(?:[eE][-+]?\d+)?', dev.split('{', 1)[1])]
        np.testing.assert_allclose(dvals, vals, rtol=0)


def test_cpu_port_rosl_convergence_and_work():
    """ROSL(11), the default method of linear models, in the host build:
    converges to the exact solution (a decade of tolerance buys about a
    decade of accuracy), needs about half the step attempts of RODAS5 at
    rtol 1e-6 and less than a quarter at rtol 1e-8, and is MORE accurate at
    the same tolerance."""
    key = cport.find_key('two_compartment_pk_model', 'central', True)
    rng = np.random.default_rng(3)
    m = ode_exact.two_compartment(True)
    for case in range(6):
        times = np.sort(rng.uniform(0.25, 24, 10))
        reg = [(float(rng.uniform(5, 15)), 0.0, 1.0, 8.0, 3)]
        psi = np.concatenate([[0.0, 0.0], np.exp(
            np.log([2.0, 1.0, 5.0, 1.0, 10.0]) + 0.4 * rng.normal(size=5))])
        pop = cport.Population(
            key, ['central.drug_concentration'], None, [times], None, None,
            [reg])
        ye, se = ode_exact.simulate(m, psi, times, reg)
        scale = np.max(np.abs(se[:, 0, :]), 0)
        scale[scale == 0] = 1.0
        errs = {}
        for rtol in (1e-6, 1e-8, 1e-10):
            y5, s5, st5, n5, r5 = pop.simulate(
                psi, rtol=rtol, atol=rtol * 0.01, stream_cb=5)
            y, s, st, n, r = pop.simulate(
                psi, rtol=rtol, atol=rtol * 0.01, stream_cb=11)
            assert st[0] == 0
            err = np.max(np.abs(y - ye[0]) / np.max(np.abs(ye[0])))
            serr = np.max(np.abs(s - se[:, 0, :]) / scale)
            err5 = np.max(np.abs(y5 - ye[0]) / np.max(np.abs(ye[0])))
            assert err < 0.1 * rtol and serr < rtol, (case, rtol, err, serr)
            assert err < 2 * err5 + 1e-13
            ratio = (n[0] + r[0]) / float(n5[0] + r5[0])
            assert ratio < (0.55 if rtol == 1e-6 else 0.3), (rtol, ratio)
            errs[rtol] = err
        assert errs[1e-10] < 1e-2 * errs[1e-6]
    # forward only (no sensitivity columns) and a subset of the columns run
    # the other two instantiations of the step (MODE 0 / MODE 2)
    y0, _, st0, _, _ = pop.simulate(
        psi, rtol=1e-9, atol=1e-11, stream_cb=11, with_sens=False)
    assert st0[0] == 0
    np.testing.assert_allclose(y0, ye[0], rtol=1e-8)
    sub = cport.Population(
        key, ['central.drug_concentration'], None, [times], None, None,
        [reg], columns=[3, 5, 0])
    _, ss, sts, _, _ = sub.simulate(psi, rtol=1e-9, atol=1e-11, stream_cb=11)
    assert sts[0] == 0
    np.testing.assert_allclose(
        ss[:, :3], se[:, 0, [3, 5, 0]], rtol=1e-7,
        atol=1e-9 * np.max(np.abs(se)))


def test_cpu_port_buffered_observation_equals_immediate_observation():
    """The ROSL kernels park (y, S) at an observation record in a one-record
    buffer and evaluate the record later (on the GPU: when every lane of the
    warp holds one).  The arithmetic is the same as observing on arrival:
    identical log-likelihoods and gradients, also with two outputs whose
    records coincide in time, and agreement with the deferred (snapshot)
    form."""
    key = cport.find_key('two_compartment_pk_model', 'central', True)
    rng = np.random.default_rng(11)
    n_ids = 6
    times, obs, out_idx, regs = [], [], [], []
    for i in range(n_ids):
        t = np.sort(np.concatenate([rng.uniform(0.25, 24, 7), [3.0, 3.0]]))
        times.append(t)
        obs.append(rng.uniform(0.2, 2.0, len(t)))
        o = rng.integers(0, 2, len(t)).astype(np.int32)
        out_idx.append(o)
        regs.append([(float(rng.uniform(5, 15)), 0.0, 1.0, 8.0, 3)])
    pop = cport.Population(
        key, ['central.drug_concentration', 'peripheral.drug_concentration'],
        ['lognormal', 'cam'], times, obs, out_idx, regs)
    X = np.column_stack([
        np.zeros((n_ids, 2)),
        np.exp(np.log([2.0, 1.0, 5.0, 1.0, 10.0])
               + 0.3 * rng.normal(size=(n_ids, 5))),
        np.full(n_ids, 0.1), np.full(n_ids, 0.05), np.full(n_ids, 0.12)])
    ids = np.arange(n_ids, dtype=np.int32)
    for with_grad in (True, False):
        a = pop.loglik(X, ids=ids, with_grad=with_grad, stream_cb=11,
                       deferred=False)
        b = pop.loglik(X, ids=ids, with_grad=with_grad, stream_cb=11,
                       buffered=True)
        c = pop.loglik(X, ids=ids, with_grad=with_grad, stream_cb=11,
                       deferred=True)
        assert not np.any(a[2]) and not np.any(b[2])
        np.testing.assert_array_equal(a[0], b[0])
        np.testing.assert_array_equal(a[1], b[1])
        np.testing.assert_array_equal(a[3], b[3])
        np.testing.assert_allclose(a[0], c[0], rtol=1e-13)
        np.testing.assert_allclose(a[1], c[1], rtol=1e-11, atol=1e-13)


def test_cpu_port_warp_synchronous_form_equals_the_general_form():
    """StreamSolver::solve_uniform (the form solve_warp_kernel runs when the
    lanes of a warp share an individual: the warp walks through the stops --
    records and dose-rate edges -- together) performs, per system, the same
    arithmetic in the same order as the general event loop: identical
    log-likelihoods, gradients and step counts, including two outputs with
    coinciding records, an invalid sigma and the forward-only solve."""
    key = cport.find_key('two_compartment_pk_model', 'central', True)
    rng = np.random.default_rng(12)
    n_ids = 8
    times, obs, out_idx, regs = [], [], [], []
    for i in range(n_ids):
        t = np.sort(np.concatenate([rng.uniform(0.25, 24, 7), [8.0, 8.0]]))
        times.append(t)
        obs.append(rng.uniform(0.2, 2.0, len(t)))
        out_idx.append(rng.integers(0, 2, len(t)).astype(np.int32))
        regs.append([(float(rng.uniform(5, 15)), 0.0, 1.0, 8.0, 3)])
    pop = cport.Population(
        key, ['central.drug_concentration', 'peripheral.drug_concentration'],
        ['lognormal', 'cam'], times, obs, out_idx, regs)
    n = 3 * n_ids
    X = np.column_stack([
        np.abs(0.01 * rng.normal(size=(n, 2))),
        np.exp(np.log([2.0, 1.0, 5.0, 1.0, 10.0])
               + 0.3 * rng.normal(size=(n, 5))),
        np.full(n, 0.1), np.full(n, 0.05), np.full(n, 0.12)])
    X[5, 7] = -1.0      # invalid sigma: -inf, nothing integrated
    ids = (np.arange(n) % n_ids).astype(np.int32)
    for method in (9, 11):
        for with_grad in (True, False):
            a = pop.loglik(X, ids=ids, with_grad=with_grad, stream_cb=method,
                           deferred=False)
            b = pop.loglik(X, ids=ids, with_grad=with_grad,
                           stream_cb=100 + method, deferred=False)
            assert a[0][5] == -np.inf and b[0][5] == -np.inf
            for u, v in zip(a, b):
                np.testing.assert_array_equal(u, v)


def test_cpu_port_split_work_items_agree_within_the_tolerance():
    """Small launches give every (system, sensitivity column) pair a work
    item of its own (SolveArgs::split, the warp-per-system form on the GPU):
    the column is integrated with the state under its own step-size control.
    Log-likelihoods and gradients agree with the all-columns solve far inside
    the integrator's tolerance; item 0 writes the log-likelihood, the
    error-model gradient and the entries of constants that are not
    integrated; an invalid sigma gives -inf / inf from every item."""
    key = cport.find_key('two_compartment_pk_model', 'central', True)
    rng = np.random.default_rng(13)
    n_ids = 10
    times = [np.sort(rng.uniform(0.25, 24, 9)) for _ in range(n_ids)]
    obs = [rng.uniform(0.2, 2.0, 9) for _ in range(n_ids)]
    regs = [[(float(rng.uniform(5, 15)), 0.0, 1.0, 8.0, 3)]
            for _ in range(n_ids)]
    pop = cport.Population(
        key, ['central.drug_concentration'], ['cam'], times, obs, None, regs)
    X = np.column_stack([
        np.abs(0.01 * rng.normal(size=(n_ids, 2))),
        np.exp(np.log([2.0, 1.0, 5.0, 1.0, 10.0])
               + 0.3 * rng.normal(size=(n_ids, 5))),
        np.full(n_ids, 0.05), np.full(n_ids, 0.12)])
    X[3, 7] = -1.0
    ids = np.arange(n_ids, dtype=np.int32)
    a = pop.loglik(X, ids=ids, stream_cb=11, deferred=False)
    b = pop.loglik(X, ids=ids, stream_cb=111, split=True, deferred=False)
    ok = np.isfinite(a[0])
    assert not ok[3] and b[0][3] == -np.inf and np.all(np.isinf(b[1][3]))
    assert not np.any(a[2]) and not np.any(b[2])
    np.testing.assert_allclose(b[0][ok], a[0][ok], rtol=2e-9)
    np.testing.assert_allclose(
        b[1][ok], a[1][ok], rtol=1e-6, atol=1e-7 * np.max(np.abs(a[1][ok])))
    # column of central.size (not integrated: output map only) and the sigmas
    assert np.all(b[1][ok][:, 2] != 0.0)
    assert np.all(b[1][ok][:, 6] == 0.0)       # peripheral.size: no effect
    assert np.all(b[1][ok][:, 7:] != 0.0)
