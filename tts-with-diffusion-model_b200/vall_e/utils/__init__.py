from .config import Config
from .utils import load_state_dict_non_strict, to_device, tree_map
