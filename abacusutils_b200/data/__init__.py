"""Device-side decoders of the Abacus particle formats (reference: abacusnbody/data/{bitpacked,pack9}.py)."""
