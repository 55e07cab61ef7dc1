"""imported by the example scripts for its side effect (environment registration) only"""
