"""The reference's models.py imports tensorflow_hub and never uses it (test infrastructure only)."""
