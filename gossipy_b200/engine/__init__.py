"""Device engine: parameter arenas, snapshots, streams, counter-based RNG, C++ scheduler."""
