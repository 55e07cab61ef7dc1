"""registers environments on import in the real package"""
