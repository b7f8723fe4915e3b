"""Inert stand-in: the reference star-imports casadi and uses nothing from it."""
