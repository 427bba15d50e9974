"""Host mirror of the reference's `src.core` package (same module and class names)."""
