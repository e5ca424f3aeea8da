"""Mirror of the reference package `src/lib` (SolutionsManagers, ReducedBasis, Estimators)."""
