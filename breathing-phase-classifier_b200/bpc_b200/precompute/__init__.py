"""Mirror of the reference package `src/precompute` (core / process / methods)."""
