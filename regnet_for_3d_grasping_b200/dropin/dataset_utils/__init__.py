"""Drop-in `dataset_utils`: only get_regiondataset is replaced; scoredataset / eval_score fall through to the reference."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
