"""Model library, mirror of ``chi.library.ModelLibrary``
(chi/library/_model_library_api.py:13-146).  Models are built from the
symbolic definitions in ``chi_b200/_library.py``; their kernels are
pre-compiled into ``libchi_b200.so``."""
from chi_b200 import _library
from chi_b200._mechanistic_models import PKPDModel, SBMLModel


class ModelLibrary(object):

    def erlotinib_tumour_growth_inhibition_model(self):
        return PKPDModel(_library.erlotinib_tumour_growth_inhibition_model())

    def one_compartment_pk_model(self):
        model = PKPDModel(_library.one_compartment_pk_model())
        model.set_outputs(['central.drug_concentration'])
        return model

    def two_compartment_pk_model(self):
        """The model of docs/source/getting_started/code/
        two_compartment_pk_model.xml (not part of chi's ModelLibrary)."""
        return PKPDModel(_library.two_compartment_pk_model())

    def tumour_growth_inhibition_model_koch(self):
        return SBMLModel(_library.tumour_growth_inhibition_model_koch())

    def tumour_growth_inhibition_model_koch_reparametrised(self):
        return SBMLModel(
            _library.tumour_growth_inhibition_model_koch_reparametrised())
