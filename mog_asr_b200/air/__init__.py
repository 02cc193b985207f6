"""PyTorch re-host of the AIR-ASR training step around the sm_100a sampler kernels
(``air/air_number_bbox_location.py`` of the reference; hyper-parameters of ``train_air_pr.py:174-238``)."""
from .model import AIRConfig, AIRModel, CudaOps, config_from_flags  # noqa: F401
from .trainer import Trainer  # noqa: F401
from .evaluate import evaluate_detection  # noqa: F401
