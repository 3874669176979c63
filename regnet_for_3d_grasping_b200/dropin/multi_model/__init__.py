"""Drop-in `multi_model` package: the reference's import paths, B200-native implementation underneath.

Sub-modules this directory does not provide (e.g. gripper_region_network.py while its restatement is pending) fall
through to the reference checkout when that is also on sys.path: the package path is extended over every
`multi_model` directory found there, this one first."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
