class DiffusionPipeline:  # plumbing stand-in; the golden driver uses its own holder object
    pass
