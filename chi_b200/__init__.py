"""chi_b200 -- B200-native likelihood engine behind chi's model API.

The public names mirror ``chi``'s for the hot path (mechanistic models, error
models, population models, log-pdfs).  Everything numerical runs in
``libchi_b200.so`` (hand-written sm_100a CUDA behind the C ABI declared in
``include/chi_b200.h``); importing the package does not need a GPU, evaluating
anything does, and nothing falls back to the CPU.
"""
from chi_b200._mechanistic_models import (  # noqa: F401
    MechanisticModel, SBMLModel, PKPDModel, ReducedMechanisticModel,
    DosingRegimen, ProtocolEvent, SimulationError, blocktrain)
from chi_b200._error_models import (  # noqa: F401
    ErrorModel, GaussianErrorModel, LogNormalErrorModel,
    ConstantAndMultiplicativeGaussianErrorModel,
    MultiplicativeGaussianErrorModel, ReducedErrorModel)
from chi_b200._covariate_models import (  # noqa: F401
    CovariateModel, LinearCovariateModel)
from chi_b200._population_models import (  # noqa: F401
    PopulationModel, PooledModel, HeterogeneousModel, GaussianModel,
    LogNormalModel, TruncatedGaussianModel, ComposedPopulationModel,
    CovariatePopulationModel)
from chi_b200._log_pdfs import (  # noqa: F401
    LogLikelihood, HierarchicalLogLikelihood, LogPosterior,
    HierarchicalLogPosterior)
from chi_b200._population_filters import (  # noqa: F401
    PopulationFilter, ComposedPopulationFilter, GaussianFilter,
    GaussianKDEFilter, GaussianMixtureFilter, LogNormalFilter,
    LogNormalKDEFilter)
from chi_b200._filter_posterior import (  # noqa: F401
    PopulationFilterLogPosterior)
from chi_b200._predictive_models import (  # noqa: F401
    PredictiveModel, PopulationPredictiveModel)
from chi_b200._averaged_predictive import (  # noqa: F401
    AveragedPredictiveModel, PosteriorPredictiveModel, PriorPredictiveModel)
from chi_b200._priors import (  # noqa: F401
    GaussianLogPrior, LogNormalLogPrior, UniformLogPrior, ComposedLogPrior)
from chi_b200._inference import (  # noqa: F401
    BatchedSamplingController, HaarioBardenetACMC, HamiltonianMCMC)
from chi_b200._problems import (  # noqa: F401
    hierarchical_log_likelihood_from_dataframe)
from chi_b200._distributed import (  # noqa: F401
    ShardedHierarchicalLogPosterior)
from chi_b200 import library  # noqa: F401
from chi_b200._lib import set_default_device  # noqa: F401

__version__ = '0.1.0'
