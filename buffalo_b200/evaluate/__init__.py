from buffalo_b200.evaluate.base import Evaluable
