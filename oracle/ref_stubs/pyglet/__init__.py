"""Inert stand-in for pyglet (rendering is out of scope)."""
image = None
