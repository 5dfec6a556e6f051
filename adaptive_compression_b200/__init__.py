"""B200-native chunked encode / select / decode path of adaptive-compression.

Same Python surface as the reference (AdaptiveCompressor, CompressionMethod plug-ins,
MarkerFinder); the work runs in hand-written sm_100a CUDA kernels behind libambc.so
(include/ambc.h).  No CPU fallback."""

__all__ = ["AdaptiveCompressor", "MarkerFinder", "CompressionMethod", "RLECompression", "DictionaryCompression",
           "HuffmanCompression", "DeltaCompression", "NoCompression"]


def __getattr__(name):  # lazy: importing the package must not need torch / a GPU
    if name == "AdaptiveCompressor":
        from .adaptive_compressor import AdaptiveCompressor
        return AdaptiveCompressor
    if name == "MarkerFinder":
        from .marker_finder import MarkerFinder
        return MarkerFinder
    if name in __all__:
        from . import compression_methods
        return getattr(compression_methods, name)
    raise AttributeError(name)
