"""Import-surface shim: `from dagr.model.networks.dagr import DAGR` etc. resolve to dagr_b200."""
