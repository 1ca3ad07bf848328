/* CAMB transfer-function reader (host, init only; SURVEY component 7).
 * Same file format handling and row selection as transfer_init.c:9-83: keep rows with
 * k > pi/BoxSize * scale, store log(k/scale) and T_nu/T_nonu (columns 1, 6, 8). */
#include <math.h>
#include <stdio.h>
#include "ksn_host.h"

/* parse one data line; returns 1 for a row of >= 8 numbers, 0 to stop, -1 for a comment */
static int camb_row(FILE *fd, double *k, double *t_nu, double *t_nonu)
{
    char line[1000];
    double c2, c3, c4, c5, c7;
    if (!fgets(line, sizeof line, fd)) return 0;
    if (line[0] == '#') return -1;
    return sscanf(line, " %lg %lg %lg %lg %lg %lg %lg %lg", k, &c2, &c3, &c4, &c5, t_nu, &c7, t_nonu) == 8;
}

void allocate_transfer_init_table(_transfer_init_table *t_init, const double BoxSize, const double UnitLength_in_cm, const double InputSpectrum_UnitLength_in_cm, const char *KspaceTransferFunction)
{
    const double scale = InputSpectrum_UnitLength_in_cm / UnitLength_in_cm;
    const double kmin = M_PI / BoxSize * scale;
    double k, tnu, tnonu;
    int st;
    FILE *fd = fopen(KspaceTransferFunction, "r");
    if (!fd) terminate(2019, "Can't read input transfer function in file '%s'\n", KspaceTransferFunction);
    t_init->NPowerTable = 0;
    while ((st = camb_row(fd, &k, &tnu, &tnonu)) != 0)
        if (st > 0 && k > kmin) t_init->NPowerTable++;
    fclose(fd);
    message(1, "Found transfer function, using %d rows. Min k used is %g.\n", t_init->NPowerTable, kmin);
    t_init->logk = mymalloc("Transfer_functions", 2 * t_init->NPowerTable * sizeof(double));
    t_init->T_nu = t_init->logk + t_init->NPowerTable;
    fd = fopen(KspaceTransferFunction, "r");
    if (!fd) terminate(2020, "Can't read input transfer function in file '%s'\n", KspaceTransferFunction);
    int count = 0;
    while (count < t_init->NPowerTable && (st = camb_row(fd, &k, &tnu, &tnonu)) != 0) {
        if (st < 0 || !(k > kmin)) continue;
        t_init->T_nu[count] = tnu / tnonu;
        t_init->logk[count] = log(k / scale);
        count++;
    }
    fclose(fd);
    if (count < t_init->NPowerTable)
        terminate(2021, "Expected %d rows in  file '%s' but only found %d\n", t_init->NPowerTable, KspaceTransferFunction, count);
}

void free_transfer_init_table(_transfer_init_table *t_init)
{
    myfree(t_init->logk);
}
