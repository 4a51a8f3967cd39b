"""Console helpers the reference imports from supplements/cli_interface.py (core.py:79)."""
from datetime import datetime


class PrintColors:
    HEADER = '\033[95m'
    BLUE = '\033[94m'
    CYAN = '\033[96m'
    GREEN = '\033[92m'
    WARNING = '\033[93m'
    FAIL = '\033[91m'
    ENDC = '\033[0m'
    BOLD = '\033[1m'
    UNDERLINE = '\033[4m'


def date_time_now():
    return datetime.now().isoformat(timespec='seconds', sep=' ')
