import os, sys
__path__ = [os.path.dirname(__file__), '/root/reference/fairseq']
__version__ = '1.0.0a0'
from fairseq.logging import meters, metrics, progress_bar
sys.modules['fairseq.meters'] = meters
sys.modules['fairseq.metrics'] = metrics
sys.modules['fairseq.progress_bar'] = progress_bar
