"""B200-native GP-surrogate + acquisition hot path behind the BoTorch Model/AcquisitionFunction surface."""
__version__ = "0.1.0"
