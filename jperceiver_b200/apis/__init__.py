"""Mirror of the reference's ``mono.apis`` (mono/apis/__init__.py:5-6)."""
from .config import Config  # noqa: F401
from .env import get_dist_info, get_root_logger, init_dist, set_random_seed  # noqa: F401
from .trainer import TrainEngine, batch_processor, build_optimizer, change_input_variable, train_mono  # noqa: F401
from .runner import Runner, StepLrPolicy, load_checkpoint, save_checkpoint  # noqa: F401,E402
