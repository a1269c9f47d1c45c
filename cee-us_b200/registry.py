"""Registry keys the new classes go under, next to the reference's own
(mbrl/controllers/__init__.py:10-14 ``valid_base_controllers``, mbrl/models/__init__.py:3-21 ``models_dict``).
INTEGRATION.md shows the two-line patch that merges them upstream."""
from .controllers import B200CuriosityMpcICem, B200MpcICem
from .models import B200GNNEnsemble, B200MLPEnsemble
from .rnd import B200RndMpcICem

CONTROLLERS = {
    "mpc-icem-b200": B200MpcICem,
    "mpc-curiosity-icem-b200": B200CuriosityMpcICem,
    "mpc-rnd-icem-b200": B200RndMpcICem,
}
FORWARD_MODELS = {
    "ParallelGNNDeterministicEnsembleB200": B200GNNEnsemble,
    "ParallelNNDeterministicEnsembleB200": B200MLPEnsemble,
}


def controller_from_string(name):
    if name not in CONTROLLERS:
        raise ImportError(f"cannot find '{name}' in known controller: {CONTROLLERS.keys()}")
    return CONTROLLERS[name]


def forward_model_from_string(name):
    if name not in FORWARD_MODELS:
        raise NotImplementedError(f"Implement model class {name} and add it to dictionary")
    return FORWARD_MODELS[name]
