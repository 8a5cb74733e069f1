"""TEST INFRASTRUCTURE ONLY - empty stand-in for librosa.display (imported, never called, on the hot path)."""
