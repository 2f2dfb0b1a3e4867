from . import squash_rewards, wrapper_base_class, torch_dataloader, nstep_return, her
from .wrapper_base_class import ReplayMemoryWrapper
from .nstep_return import NStepReturn
from .her import HindsightNStepReplay, SampleTimeHindsight
from .squash_rewards import SquashRewards
from .torch_dataloader import TorchDataLoader, ConfigurationError
