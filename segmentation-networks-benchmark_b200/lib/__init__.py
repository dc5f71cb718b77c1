"""Host-side mirror of the reference's `lib` package for the tiled-inference hot path (same module names)."""
