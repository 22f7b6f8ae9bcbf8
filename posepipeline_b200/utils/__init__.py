"""Drop-ins for the reference helpers that sit right next to the hot path."""
