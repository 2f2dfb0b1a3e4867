from . import squash_rewards, wrapper_base_class, torch_dataloader, nstep_return, her, her_vmap
from . import nstep_return as nstep_return_vmap  # the reference keeps NStepReturnVmap in its own module
from .wrapper_base_class import ReplayMemoryWrapper
from .nstep_return import NStepReturn, NStepReturnVmap
from .her_vmap import HindsightVmapWrite, HindsightVmapRead
from .her import HindsightNStepReplay, SampleTimeHindsight, IgnoreKeys
from .squash_rewards import SquashRewards
from .torch_dataloader import TorchDataLoader, ConfigurationError
