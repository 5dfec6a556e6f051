"""MarkerFinder backed by the CUDA n-gram presence bitmap (libambc.so).
Mirror of the reference's class (marker_finder.py:6-123)."""
import time

from . import engine


class MarkerFinder:
    def __init__(self, max_marker_length=32):
        self.max_marker_length = max_marker_length

    def find_marker(self, file_data, sample_size=None):
        """Shortest bit string (smallest value first) that does not occur in the MSB-first bit
        stream of file_data -> (marker bytes left-aligned, length in bits).  With sample_size the
        search runs on every (len // sample_size)-th byte, concatenated and cut to sample_size
        bytes, exactly like marker_finder.py:38-48.  Raises ValueError when nothing up to
        max_marker_length bits is absent (:123)."""
        start = time.time()
        if sample_size and len(file_data) > sample_size:
            step = len(file_data) // sample_size
            file_data = bytes(file_data[0:len(file_data):step][:sample_size])
        t = engine.to_device(file_data)
        marker, length = engine.find_marker_device(t, min(int(self.max_marker_length), 32))
        self.last_elapsed = time.time() - start
        return marker, length
