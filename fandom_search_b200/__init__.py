"""B200-native reuse-search hot path of senderle/fandom-search (drop-in for search.py)."""
__all__ = ["build", "engine"]
