"""frankenstein-b200: the neural-encoder training step of ALVI-Labs/frankenstein on sm_100a kernels."""
__version__ = "0.1.0"
