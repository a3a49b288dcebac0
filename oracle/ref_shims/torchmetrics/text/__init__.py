class Perplexity:          # placeholder: importing esme.variant must work, predict_pseudoperplexity is not run
    def __init__(self, *a, **kw):
        raise NotImplementedError('torchmetrics is not installed; this is an import stub')
