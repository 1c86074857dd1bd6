"""Mirrors of the reference's `utils/` pieces that sit next to the hot path (SURVEY §8f "next" rows)."""
