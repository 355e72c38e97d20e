"""Built-in symbolic definitions of the models in chi's model library.

chi ships its library models as SBML files
(chi/library/model_library/*.xml, chi/library/_model_library_api.py:13-146;
the two-compartment model lives in
docs/source/getting_started/code/two_compartment_pk_model.xml).  Those files
are inputs of the reference and are not redistributed here; the same models
are defined programmatically below so that the pre-compiled kernels exist on a
machine that has neither the SBML files nor a compiler.  Reading the SBML file
with :func:`chi_b200._sbml.read_sbml` yields the same model key
(tests/test_host_models.py checks this whenever the reference is present).
"""
import sympy as sp

from chi_b200._sbml import SymbolicModel


def _build(name, compartments, species, parameters, rates, time_unit='day'):
    """compartments: [name]; species: [(compartment, name)];
    parameters: [name] (global constants); rates: {state qname: str expr}
    written with '__' instead of '.' in names."""
    m = SymbolicModel()
    m.name = name
    m.time_unit = time_unit
    for c in compartments:
        m.constants[c + '.size'] = 1.0
    for comp, s in species:
        amount = '%s.%s_amount' % (comp, s)
        m.intermediates['%s.%s_concentration' % (comp, s)] = \
            m.symbol(amount) / m.symbol(comp + '.size')
        m.initial_values[amount] = 0.0
    for p in parameters:
        m.constants['global.' + p] = 1.0
    names = set(m.constants) | set(m.intermediates) | set(rates) | \
        set(m.initial_values)
    local = {n.replace('.', '__'): m.symbol(n) for n in names}
    for state, expr in rates.items():
        if state in m.constants:
            m.initial_values[state] = m.constants.pop(state)
        m.states.append(state)
        m.initial_values.setdefault(state, 0.0)
        m.rhs[state] = sp.sympify(expr, locals=local)
    return m


def one_compartment_pk_model():
    """`ModelLibrary.one_compartment_pk_model`
    (chi/library/_model_library_api.py:43-67; pk_one_comp.xml:24-52)."""
    return _build(
        'one_compartment_pk_model', ['central'], [('central', 'drug')],
        ['elimination_rate'],
        {'central.drug_amount':
            '-central__size * global__elimination_rate '
            '* central__drug_concentration'})


def two_compartment_pk_model():
    """docs/source/getting_started/code/two_compartment_pk_model.xml:24-93
    (used by docs/source/getting_started/code/3_fitting_models_1.py:62)."""
    el = 'central__size * global__elimination_rate ' \
        '* central__drug_concentration'
    c2p = 'central__size * global__transition_rate_c2p ' \
        '* central__drug_concentration'
    p2c = 'peripheral__size * global__transition_rate_p2c ' \
        '* peripheral__drug_concentration'
    return _build(
        'two_compartment_pk_model', ['central', 'peripheral'],
        [('central', 'drug'), ('peripheral', 'drug')],
        ['elimination_rate', 'transition_rate_c2p', 'transition_rate_p2c'],
        {'central.drug_amount': '-(%s) - (%s) + (%s)' % (el, c2p, p2c),
         'peripheral.drug_amount': '(%s) - (%s)' % (c2p, p2c)})


def erlotinib_tumour_growth_inhibition_model():
    """`ModelLibrary.erlotinib_tumour_growth_inhibition_model`
    (chi/library/_model_library_api.py:33-41;
    temporary_full_pkpd_model.xml:41-106)."""
    return _build(
        'erlotinib_tumour_growth_inhibition_model', ['central'],
        [('central', 'drug')],
        ['critical_volume', 'elimination_rate', 'kappa', 'lambda',
         'tumour_volume'],
        {'central.drug_amount':
            '-central__size * global__elimination_rate '
            '* central__drug_concentration',
         'global.tumour_volume':
            'global__lambda * global__critical_volume '
            '* global__tumour_volume '
            '/ (global__tumour_volume + global__critical_volume) '
            '- global__kappa * central__drug_concentration '
            '* global__tumour_volume'})


def tumour_growth_inhibition_model_koch():
    """`ModelLibrary.tumour_growth_inhibition_model_koch`
    (chi/library/_model_library_api.py:69-104; tgi_Koch_2009.xml:40-86)."""
    return _build(
        'tumour_growth_inhibition_model_koch', [], [],
        ['tumour_volume', 'lambda_0', 'lambda_1', 'kappa',
         'drug_concentration'],
        {'global.tumour_volume':
            '2 * global__lambda_0 * global__lambda_1 * global__tumour_volume '
            '/ (2 * global__lambda_0 * global__tumour_volume '
            '+ global__lambda_1) '
            '- global__kappa * global__drug_concentration '
            '* global__tumour_volume'})


def tumour_growth_inhibition_model_koch_reparametrised():
    """`ModelLibrary.tumour_growth_inhibition_model_koch_reparametrised`
    (chi/library/_model_library_api.py:106-146;
    tgi_Koch_2009_reparametrised.xml)."""
    return _build(
        'tumour_growth_inhibition_model_koch_reparametrised', [], [],
        ['tumour_volume', 'lambda', 'critical_volume', 'kappa',
         'drug_concentration'],
        {'global.tumour_volume':
            'global__lambda * global__critical_volume * global__tumour_volume '
            '/ (global__tumour_volume + global__critical_volume) '
            '- global__kappa * global__drug_concentration '
            '* global__tumour_volume'})


# name -> (constructor, compartments that can be dosed)
BUILTIN = {
    'one_compartment_pk_model': (one_compartment_pk_model, ['central']),
    'two_compartment_pk_model': (two_compartment_pk_model, ['central']),
    'erlotinib_tumour_growth_inhibition_model': (
        erlotinib_tumour_growth_inhibition_model, ['central']),
    'tumour_growth_inhibition_model_koch': (
        tumour_growth_inhibition_model_koch, []),
    'tumour_growth_inhibition_model_koch_reparametrised': (
        tumour_growth_inhibition_model_koch_reparametrised, []),
}


def builtin_variants():
    """Yields every (name, administration, SymbolicModel) that is compiled
    into the core library ahead of time."""
    for name, (ctor, compartments) in BUILTIN.items():
        yield name, None, ctor()
        for comp in compartments:
            m = ctor()
            m.add_dose_rate(comp + '.drug_amount')
            yield name, (comp, True), m
            m = ctor()
            dose = m.add_dose_compartment(comp + '.drug_amount')
            m.add_dose_rate(dose)
            yield name, (comp, False), m
