/* oracle/orc_luts.c -- TEST INFRASTRUCTURE. EV look-up tables, restating main.c:128-196. */
#include <math.h>
#include <pthread.h>
#include <string.h>
#include "oracle.h"

static int    g_raw2ev[16384 + ORC_MAX_BLACK];
static double g_raw2evf[16384 + ORC_MAX_BLACK];
static int    g_ev2raw[24 * ORC_EV_RES];
static pthread_once_t g_once = PTHREAD_ONCE_INIT;

static void build(void)
{
    /* main.c:134-141 / 160-167: entries below black stay 0; entry d = v-black holds log2(d)*EV.
       d == 0 gives -inf: the int cast of -inf is INT_MIN on x86 (cvttsd2si), kept deliberately. */
    memset(g_raw2ev, 0, sizeof(g_raw2ev));
    memset(g_raw2evf, 0, sizeof(g_raw2evf));
    for (int d = 0; d < 16384; d++) {
        double e = log2((double)d) * ORC_EV_RES;
        g_raw2evf[d + ORC_MAX_BLACK] = e;
        g_raw2ev[d + ORC_MAX_BLACK] = (d == 0) ? INT32_MIN : (int)e;
    }
    /* main.c:187-192: float division, double pow, truncating cast */
    for (int e = -10 * ORC_EV_RES; e < 14 * ORC_EV_RES; e++)
        g_ev2raw[e + 10 * ORC_EV_RES] = (int)pow(2, (float)e / ORC_EV_RES);
}

const int *orc_raw2ev(int black)
{
    pthread_once(&g_once, build);
    if (black > ORC_MAX_BLACK) return NULL;           /* main.c:170-174 */
    return &g_raw2ev[ORC_MAX_BLACK - black];
}

const double *orc_raw2evf(int black)
{
    pthread_once(&g_once, build);
    if (black > ORC_MAX_BLACK) return NULL;
    return &g_raw2evf[ORC_MAX_BLACK - black];
}

const int *orc_ev2raw(void)
{
    pthread_once(&g_once, build);
    return g_ev2raw + 10 * ORC_EV_RES;
}
