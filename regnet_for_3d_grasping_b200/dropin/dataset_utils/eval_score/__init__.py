"""Drop-in `dataset_utils.eval_score`: eval.py is replaced (batched view-collision filter), configs / eval_utils fall
through to the reference."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
