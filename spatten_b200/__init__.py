"""spatten_b200: B200-native sparse-attention decode path behind the spatten_llm API."""
__version__ = "0.1.0"
