def __getattr__(name):
    raise RuntimeError("matplotlib stub: image.%s is not available" % name)
