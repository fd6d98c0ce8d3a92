/*
 * dpe_flow.h -- C handle on the host-side flow mirror (libdpe_flow.so): the reference's
 * console + FlowMgr + DPEFlow + modules for the DPE path (cudarecv/cudarecv/src/main.cu,
 * cmdFlow.cpp:16-166, dsp/src/flowmgr.cpp, dsp/src/dpeflow.cpp), re-implemented in
 * navlab-dpe-sdr_b200/host/ on top of include/dpe_b200.h.  Used by the parity tests and by
 * anyone who wants the `newflow dpe / loadflow / startflow` surface without a tty.
 */
#ifndef DPE_FLOW_H_
#define DPE_FLOW_H_
#ifdef __cplusplus
extern "C" {
#endif

typedef struct dpe_shell dpe_shell;

dpe_shell* dpe_shell_create(void);
void dpe_shell_destroy(dpe_shell* sh);
/* one console command line; 0 ok, -1 error, 1 quit requested (console/src/cmdParser.cpp:102) */
int dpe_shell_exec(dpe_shell* sh, const char* line);
/* start the modules and step the flow on the caller's thread for at most max_epochs epochs
 * (-1 = until a module ends the flow); the synchronous form of `startflow` + `waitflow` */
int dpe_shell_run_blocking(dpe_shell* sh, const char* flow, long max_epochs);
/* {runCount, avg_us, min_us, max_us, total_s}: the statistics Flow::FlowThread prints (flow.cu:172-191) */
int dpe_shell_flow_stats(dpe_shell* sh, const char* flow, double* out5);
/* copy a HOST output port as doubles; returns the element count or -1 */
long dpe_shell_read_port(dpe_shell* sh, const char* flow, const char* module, const char* port, double* out, long cap);

/* host GPS code exposed for parity tests: CHM_Get_Sat_Pos (cuchanmgr.cu:85-210) with the nearest-TOE
 * ephemeris of a RINEX 2.x file, and BCM_InitPosGrid (batchcorrmanifold.cu:148-255) */
int dpe_host_sat_position(const char* rinex_path, int prn, double tx_time, double* state8);
int dpe_host_make_grid(const int* dims4, const double* spacing4, int grid_type, double* out, long cap);
/* the file readers either side of the path, for CPU parity tests: the handoff CSV
 * (dpinit.cpp:247-400) flattened as [n, rxTime, bytes_read, t_oe, X_ECEF(8), then 8 rows of n:
 * prn, rc, ri, fc, fi, cp, cp_timestamp, TOW]; the grid CSV (batchcorrmanifold.cu:2433-2444) as
 * [G][4].  Return the element / candidate count or -1.                                         */
long dpe_host_read_handoff(const char* path, double* out, long cap);
long dpe_host_read_grid(const char* path, double* out, long cap);

#ifdef __cplusplus
}
#endif
#endif
