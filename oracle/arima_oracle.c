/* ARIMA oracle (C port) -- placeholder translation unit; filled in by the ARIMA milestone.
 * TEST INFRASTRUCTURE ONLY (see tad_oracle.c). */
int tad_oracle_arima_available(void) { return 0; }
