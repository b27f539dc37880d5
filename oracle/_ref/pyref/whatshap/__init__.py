"""Stub package (test infrastructure): only the compiled reference binding whatshap.core lives here."""
