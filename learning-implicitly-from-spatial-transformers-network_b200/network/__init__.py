"""Mirror of the reference's `network` package for the LIST hot path (modules, models, executors)."""
