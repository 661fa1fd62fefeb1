"""B200-native Whitted hot path (see rt_b200.py; the directory name is not importable, add it to sys.path)."""
