from buffalo_b200.misc import aux, log
from buffalo_b200.misc.log import get_log_level, get_logger, set_log_level

_aux = aux
util = aux  # the reference ships misc/util.py as a duplicate of misc/_aux.py
