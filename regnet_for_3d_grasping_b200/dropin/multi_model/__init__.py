"""Drop-in `multi_model` package: same import paths as the reference, B200-native implementation underneath."""
