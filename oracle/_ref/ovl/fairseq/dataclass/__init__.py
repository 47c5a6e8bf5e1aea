import os
__path__ = [os.path.dirname(__file__), '/root/reference/fairseq/dataclass']
from .configs import FairseqDataclass
from .constants import ChoiceEnum
