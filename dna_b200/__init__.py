"""hyena-b200: B200-native (sm_100a) kernels for the HyenaDNA hot path behind the reference's
HyenaOperator / fftconv_func call surface."""
__version__ = "0.1.0"
