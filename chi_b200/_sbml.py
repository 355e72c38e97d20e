"""Minimal SBML (L3 core) reader -> symbolic ODE model.

Replaces, for the likelihood hot path, the pieces of myokit that chi uses to
turn an SBML file into a simulator:

* ``myokit.formats.sbml.SBMLImporter().model(file)``
  (call site chi/_mechanistic_models.py:136),
* the model-editing calls chi makes to add a dose compartment / dose rate
  (chi/_mechanistic_models.py:563-636),
* myokit's symbolic derivation of the sensitivity equations.

Only the SBML constructs used by the reference's model library and docs are
supported (compartments, species, parameters, reactions with kineticLaw,
rateRule, assignmentRule, initialAssignment-free); anything else raises
``NotImplementedError`` loudly.

Naming follows myokit's importer as pinned by chi's tests
(chi/tests/test_mechanistic_models.py:127-142, 451-481): compartment ``c``
becomes component ``c`` with constant ``c.size``; species ``s`` in ``c``
becomes state ``c.s_amount`` and intermediary ``c.s_concentration``; global
parameters live in component ``global``.
"""
import xml.etree.ElementTree as ET

import sympy as sp

_MATHML = '{http://www.w3.org/1998/Math/MathML}'


def _strip(tag):
    return tag.split('}', 1)[1] if '}' in tag else tag


class SymbolicModel(object):
    """ODE model in symbolic form.

    Attributes
    ----------
    states : list[str]              qualified names, definition order
    rhs : dict[str, sympy.Expr]     d(state)/dt in terms of symbols()
    constants : dict[str, float]    literal constants and default values
    intermediates : dict[str, sympy.Expr]
    initial_values : dict[str, float]
    dose_rate : sympy.Symbol or None   symbol of the pace-bound dose rate
    time_unit : str or None
    """

    def __init__(self):
        self.states = []
        self.rhs = {}
        self.constants = {}
        self.intermediates = {}
        self.initial_values = {}
        self.dose_rate = None
        self.dosed_state = None
        self.time_unit = None
        self.name = ''
        self._symbols = {}

    # -- symbols ---------------------------------------------------------
    def symbol(self, qname):
        if qname not in self._symbols:
            self._symbols[qname] = sp.Symbol(
                qname.replace('.', '__'), real=True)
        return self._symbols[qname]

    def components(self):
        names = list(self.states) + list(self.constants) + \
            list(self.intermediates)
        return sorted(set(n.split('.', 1)[0] for n in names))

    def has_variable(self, qname):
        return (qname in self.states or qname in self.constants
                or qname in self.intermediates)

    def clone(self):
        m = SymbolicModel()
        m.states = list(self.states)
        m.rhs = dict(self.rhs)
        m.constants = dict(self.constants)
        m.intermediates = dict(self.intermediates)
        m.initial_values = dict(self.initial_values)
        m.dose_rate = self.dose_rate
        m.dosed_state = self.dosed_state
        m.time_unit = self.time_unit
        m.name = self.name
        m._symbols = dict(self._symbols)
        return m

    # -- chi's structural edits -------------------------------------------
    def add_dose_rate(self, state):
        """chi/_mechanistic_models.py:611-636 (`_add_dose_rate`)."""
        comp = state.split('.', 1)[0]
        self.dose_rate = self.symbol(comp + '.dose_rate')
        self.dosed_state = state
        self.rhs[state] = self.rhs[state] + self.dose_rate

    def add_dose_compartment(self, state):
        """chi/_mechanistic_models.py:563-609 (`_add_dose_compartment`).
        Returns the name of the new dose-compartment state."""
        comp = 'dose'
        k = 1
        while comp in self.components():
            comp = 'dose_%d' % k
            k += 1
        amount = comp + '.drug_amount'
        ka = comp + '.absorption_rate'
        self.states.append(amount)
        self.initial_values[amount] = 0.0
        self.constants[ka] = 1.0
        self.rhs[amount] = -self.symbol(ka) * self.symbol(amount)
        self.rhs[state] = self.rhs[state] + \
            self.symbol(ka) * self.symbol(amount)
        return amount

    # -- derived quantities -------------------------------------------------
    def expanded(self, expr):
        """Substitutes intermediary variables until only states, constants
        and the dose rate remain."""
        inter = {self.symbol(k): v for k, v in self.intermediates.items()}
        for _ in range(len(inter) + 1):
            new = expr.xreplace(inter)
            if new == expr:
                return new
            expr = new
        raise ValueError('Cyclic assignment rules.')

    def expression(self, qname):
        """Expression of a state/intermediary/constant in base symbols."""
        if qname in self.intermediates:
            return self.expanded(self.intermediates[qname])
        if qname in self.states or qname in self.constants:
            return self.symbol(qname)
        raise KeyError(qname)


def _mathml(node, resolve):
    tag = _strip(node.tag)
    if tag == 'math':
        children = list(node)
        if len(children) != 1:
            raise NotImplementedError('math element with %d children'
                                      % len(children))
        return _mathml(children[0], resolve)
    if tag == 'ci':
        return resolve(node.text.strip())
    if tag == 'cn':
        kind = node.attrib.get('type', 'real')
        if kind == 'e-notation':
            parts = [t.strip() for t in node.itertext() if t.strip()]
            return sp.Float(float(parts[0]) * 10 ** float(parts[1]))
        if kind == 'rational':
            parts = [t.strip() for t in node.itertext() if t.strip()]
            return sp.Rational(int(parts[0]), int(parts[1]))
        text = node.text.strip()
        try:
            return sp.Integer(int(text))
        except ValueError:
            return sp.Float(float(text))
    if tag == 'apply':
        children = list(node)
        op = _strip(children[0].tag)
        args = [_mathml(c, resolve) for c in children[1:]]
        if op == 'times':
            return sp.Mul(*args)
        if op == 'plus':
            return sp.Add(*args)
        if op == 'minus':
            if len(args) == 1:
                return -args[0]
            return args[0] - args[1]
        if op == 'divide':
            return args[0] / args[1]
        if op == 'power':
            return args[0] ** args[1]
        if op == 'exp':
            return sp.exp(args[0])
        if op == 'ln':
            return sp.log(args[0])
        if op == 'root':
            return sp.sqrt(args[0])
        if op == 'abs':
            return sp.Abs(args[0])
        raise NotImplementedError('MathML operator <%s>' % op)
    if tag == 'exponentiale':
        return sp.E
    if tag == 'pi':
        return sp.pi
    raise NotImplementedError('MathML element <%s>' % tag)


def read_sbml(path):
    """Parses an SBML file into a :class:`SymbolicModel`."""
    tree = ET.parse(path)
    root = tree.getroot()
    model_el = None
    for child in root:
        if _strip(child.tag) == 'model':
            model_el = child
    if model_el is None:
        raise ValueError('No <model> element in ' + str(path))

    m = SymbolicModel()
    m.name = model_el.attrib.get('id', '')
    m.time_unit = model_el.attrib.get('timeUnits')

    def section(name):
        for child in model_el:
            if _strip(child.tag) == name:
                return list(child)
        return []

    for name in ('listOfFunctionDefinitions', 'listOfEvents',
                 'listOfConstraints', 'listOfInitialAssignments'):
        if section(name):
            raise NotImplementedError('SBML ' + name + ' is not supported.')

    sid = {}            # SBML id -> sympy expression used inside formulas
    compartments = {}
    for c in section('listOfCompartments'):
        cid = c.attrib['id']
        cname = c.attrib.get('name', cid)
        q = cname + '.size'
        m.constants[q] = float(c.attrib.get('size', 1.0))
        compartments[cid] = cname
        sid[cid] = m.symbol(q)

    species = {}        # SBML id -> amount qname
    boundary = set()
    for s in section('listOfSpecies'):
        s_id = s.attrib['id']
        s_name = s.attrib.get('name', s_id)
        comp = compartments[s.attrib['compartment']]
        amount = '%s.%s_amount' % (comp, s_name)
        conc = '%s.%s_concentration' % (comp, s_name)
        size = m.symbol(comp + '.size')
        if 'initialAmount' in s.attrib:
            init = float(s.attrib['initialAmount'])
        elif 'initialConcentration' in s.attrib:
            init = float(s.attrib['initialConcentration']) * \
                m.constants[comp + '.size']
        else:
            init = 0.0
        species[s_id] = amount
        m.initial_values[amount] = init
        m.intermediates[conc] = m.symbol(amount) / size
        if s.attrib.get('hasOnlySubstanceUnits', 'false') == 'true':
            sid[s_id] = m.symbol(amount)
        else:
            sid[s_id] = m.symbol(conc)
        if s.attrib.get('boundaryCondition', 'false') == 'true':
            boundary.add(s_id)

    param_q = {}
    for p in section('listOfParameters'):
        p_id = p.attrib['id']
        q = 'global.' + p_id
        param_q[p_id] = q
        m.constants[q] = float(p.attrib.get('value', 0.0))
        sid[p_id] = m.symbol(q)

    def resolve(name):
        if name not in sid:
            raise KeyError('Unknown identifier <%s> in SBML math.' % name)
        return sid[name]

    # reactions -> species amount rates
    rates = {}
    for r in section('listOfReactions'):
        law = None
        reactants, products = [], []
        for child in r:
            tag = _strip(child.tag)
            if tag == 'kineticLaw':
                for sub in child:
                    if _strip(sub.tag) == 'listOfLocalParameters' and \
                            len(list(sub)):
                        raise NotImplementedError(
                            'Local kinetic-law parameters.')
                    if _strip(sub.tag) == 'math':
                        law = _mathml(sub, resolve)
            elif tag in ('listOfReactants', 'listOfProducts'):
                for ref in child:
                    st = float(ref.attrib.get('stoichiometry', 1.0))
                    st = sp.Integer(int(st)) if st == int(st) else sp.Float(st)
                    (reactants if tag == 'listOfReactants'
                     else products).append((ref.attrib['species'], st))
        if law is None:
            continue
        for s_id, st in reactants:
            if s_id not in boundary:
                rates[s_id] = rates.get(s_id, 0) - st * law
        for s_id, st in products:
            if s_id not in boundary:
                rates[s_id] = rates.get(s_id, 0) + st * law

    for s_id, amount in species.items():
        if s_id in rates:
            m.states.append(amount)
            m.rhs[amount] = rates[s_id]
        else:
            # species that never changes: a literal constant
            m.constants[amount] = m.initial_values.pop(amount)

    for rule in section('listOfRules'):
        tag = _strip(rule.tag)
        var = rule.attrib['variable']
        math = None
        for sub in rule:
            if _strip(sub.tag) == 'math':
                math = _mathml(sub, resolve)
        if var in param_q:
            q = param_q[var]
        elif var in species:
            q = species[var]
        elif var in compartments:
            q = compartments[var] + '.size'
        else:
            raise KeyError('Rule for unknown variable <%s>.' % var)
        if tag == 'rateRule':
            value = m.constants.pop(q, m.initial_values.get(q, 0.0))
            if q not in m.states:
                m.states.append(q)
            m.initial_values[q] = value
            m.rhs[q] = math
        elif tag == 'assignmentRule':
            m.constants.pop(q, None)
            m.intermediates[q] = math
        else:
            raise NotImplementedError('SBML rule <%s>' % tag)

    return m
