"""Host-side mirrors of the reference's `utils/` helpers that sit on the loader side of the hot path."""
