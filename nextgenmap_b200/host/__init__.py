from .cuda_sw import CudaSW, Align, Params, load_library, NgmB200Error, EncodedReference, PrefixTableFile, CsParams, PeParams  # noqa: F401
