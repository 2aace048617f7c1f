/* ARIMA has no C port: the authoritative ARIMA oracle is oracle/arima_oracle.py (SciPy's own L-BFGS-B and Brent,
 * i.e. the routines statsmodels / scipy.stats.boxcox call).  This translation unit only keeps the shared library's
 * symbol table stable.  TEST INFRASTRUCTURE ONLY (see tad_oracle.c). */
int tad_oracle_arima_available(void) { return 0; }
