/*
 * apm_files.c -- the text files that connect the phases and that `analyse` and the
 * reference's post-processing tools read.  Formats are byte-compatible with the
 * reference; each writer cites the code it matches.
 *
 * With N_ENSEMBLES > 1 every ensemble is one complete copy of the reference's
 * working directory below ens<e>/ (the reference's own advice for independent
 * runs, doc/faq.rst:45-46); `params` and `data` are shared from the current
 * directory.  apm_set_output_dir() selects where the functions below read/write.
 */
#include <errno.h>
#include <sys/stat.h>
#include <sys/types.h>
#include "apm_host.h"
#include "apm_session.h"

static char out_dir[64] = "";

void apm_set_output_dir(int ensemble) {
	if (N_ENSEMBLES == 1 || ensemble < 0) {
		out_dir[0] = 0;
		return;
	}
	snprintf(out_dir, sizeof(out_dir), "ens%d/", ensemble);
	if (mkdir(out_dir, 0777) != 0 && errno != EEXIST) {
		perror("could not create the ensemble directory");
		exit(1);
	}
}

const char * apm_out_path(const char * name) {
	static __thread char buf[4][APM_PATH_MAX + 64]; /* per thread: analyse reads its files in parallel */
	static __thread int slot = 0;
	char * b = buf[slot++ & 3];
	snprintf(b, sizeof(buf[0]), "%s%s", out_dir, name);
	return b;
}

/* ---- params_suggested: best min max name step (reference
 * src/parallel_tempering_config.c:28-47) ------------------------------------------ */
void write_params_file(mcmc * m) {
	const char * path = apm_out_path(PARAMS_FILENAME "_suggested");
	FILE * f = fopen(path, "w");
	unsigned int i;
	if (f == NULL) {
		fprintf(stderr, "Could not write to file %s\n", path);
		return;
	}
	for (i = 0; i < get_n_par(m); i++)
		fprintf(f, DUMP_FORMAT "\t" DUMP_FORMAT "\t" DUMP_FORMAT "\t%s\t" DUMP_FORMAT "\n",
				gsl_vector_get(get_params_best(m), i), get_params_min_for(m, i),
				get_params_max_for(m, i), get_params_descr(m)[i], get_steps_for(m, i));
	fclose(f);
	printf("new suggested parameters file has been written\n");
}

/* ---- calibration_summary (reference src/parallel_tempering_config.c:49-94) -------- */
void write_calibration_summary(mcmc ** chains, unsigned int n_chains) {
	const double beta_0 = get_beta(chains[n_chains - 1]);
	const unsigned int n_par = get_n_par(chains[0]);
	unsigned int i, j;
	FILE * f = fopen(apm_out_path("calibration_summary"), "w");
	if (f == NULL) {
		fprintf(stderr, "Could not write to file calibration_summary\n");
		return;
	}
	fprintf(f, "Summary of calibrations\n\nBETA TABLE\nChain # | Calculated | Calibrated\n");
	for (i = 0; i < n_chains; i++)
		fprintf(f, "Chain %d | " DUMP_FORMAT " | %f\n", i, get_chain_beta(i, n_chains, beta_0),
				get_beta(chains[i]));
	fprintf(f, "\nSTEPWIDTH TABLE\nChain # | Calibrated stepwidths... \n");
	for (i = 0; i < n_chains; i++) {
		fprintf(f, "%d", i);
		for (j = 0; j < n_par; j++)
			fprintf(f, "\t" DUMP_FORMAT, get_steps_for(chains[i], j));
		fprintf(f, "\n");
	}
	fprintf(f, "\nSTEPWIDTH ESTIMATE TABLE\n"
			"If you find that the estimate deviates much or systematically from the "
			"calibrated stepwidths, please notify the authors.\n"
			"Chain # | Calculated stepwidths... \n");
	for (i = 0; i < n_chains; i++) {
		const double scale = pow(get_beta(chains[i]), -0.5);
		fprintf(f, "%d", i);
		for (j = 0; j < n_par; j++)
			fprintf(f, "\t" DUMP_FORMAT, get_steps_for(chains[0], j) * scale);
		fprintf(f, "\n");
	}
	fclose(f);
	printf("calibration summary has been written\n");
}

/* ---- calibration_results: one line per chain, beta, steps, params, each "%.15e"
 * (reference src/parallel_tempering_config.c:130-202) -------------------------------- */
void read_calibration_file(mcmc ** chains, unsigned int n_chains) {
	const unsigned int n_par = get_n_par(chains[0]);
	const char * path = apm_out_path(CALIBRATION_FILE);
	FILE * f = fopen(path, "r");
	unsigned int i, j;
	double v;
	if (f == NULL) {
		fprintf(stderr, "could not read calibration file '%s': %s\n", path, strerror(errno));
		exit(1);
	}
	for (i = 0; i < n_chains; i++) {
		int bad = fscanf(f, "%lf", &v) != 1;
		if (!bad)
			set_beta(chains[i], v);
		for (j = 0; j < n_par && !bad; j++) {
			bad = fscanf(f, "%lf", &v) != 1;
			if (!bad)
				set_steps_for(chains[i], v, j);
		}
		for (j = 0; j < n_par && !bad; j++) {
			bad = fscanf(f, "%lf", &v) != 1;
			if (!bad)
				set_params_for(chains[i], v, j);
		}
		if (bad) {
			fprintf(stderr, "could not read %d chain calibrations. \nError with line %d.\n", n_chains,
					i + 1);
			exit(1);
		}
		set_params_best(chains[i], get_params(chains[i]));
	}
	fclose(f);
}

void write_calibrations_file(mcmc ** chains, const unsigned int n_chains) {
	const unsigned int n_par = get_n_par(chains[0]);
	const char * path = apm_out_path(CALIBRATION_FILE);
	FILE * f = fopen(path, "w");
	unsigned int i, j;
	if (f == NULL) {
		perror("error writing to calibration results file");
		exit(1);
	}
	for (j = 0; j < n_chains; j++) {
		fprintf(f, DUMP_FORMAT, get_beta(chains[j]));
		for (i = 0; i < n_par; i++)
			fprintf(f, "\t" DUMP_FORMAT, get_steps_for(chains[j], i));
		for (i = 0; i < n_par; i++)
			fprintf(f, "\t" DUMP_FORMAT, get_params_for(chains[j], i));
		fprintf(f, "\n");
	}
	if (fclose(f) != 0) {
		perror("error writing to calibration results file");
		exit(1);
	}
	printf("wrote calibration results for %d chains to %s\n", n_chains, path);
}

/* ---- calibration_progress.data (reference src/markov_chain_calibrate.c:1053,1143-1146).
 * The reference reopens the file with "w" for every chain it calibrates, so after a phase
 * it holds the rows of the last chain calibrated; same here. -------------------------- */
void apm_write_calibration_progress(const apm_gpu_calib_progress * rows, long long n, int chain) {
	FILE * f = fopen(apm_out_path("calibration_progress.data"), "w");
	long long r;
	if (f == NULL)
		return;
	for (r = 0; r < n; r++)
		if (rows[r].chain == chain)
			fprintf(f, "%d\t%lu\t%f\t%f\t%f\n", rows[r].param, (unsigned long) rows[r].iter,
					rows[r].step_normalised, rows[r].accept_rate, -1.);
	fclose(f);
}

/* ---- per-parameter traces <name><suffix>-<index>.prob.dump, one "%.15e" per line
 * (reference src/mcmc_dump.c:60-113) ---------------------------------------------------- */
void mcmc_open_dump_files(mcmc * m, const char * suffix, int index, char * mode) {
	unsigned int i;
	char name[APM_PATH_MAX];
	m->files = (FILE **) calloc(m->n_par, sizeof(FILE *));
	assert(m->files != NULL);
#ifdef NODUMP
	return;
#endif
	for (i = 0; i < get_n_par(m); i++) {
		snprintf(name, sizeof(name), "%s%s-%d.prob.dump", m->params_descr[i], suffix, index);
		m->files[i] = fopen(apm_out_path(name), mode);
		if (m->files[i] == NULL) {
			fprintf(stderr, "opening file %s failed\n", name);
			exit(1);
		}
		setvbuf(m->files[i], NULL, _IOFBF, 1 << 20);
	}
}

void mcmc_dump_current(const mcmc * m) {
	unsigned int i;
	if (m->files == NULL)
		return;
	for (i = 0; i < get_n_par(m); i++)
		if (m->files[i] != NULL)
			fprintf(m->files[i], DUMP_FORMAT "\n", gsl_vector_get(m->params, i));
}

void mcmc_dump_flush(const mcmc * m) {
	unsigned int i;
	if (m->files == NULL)
		return;
	for (i = 0; i < get_n_par(m); i++)
		if (m->files[i] != NULL)
			fflush(m->files[i]);
}

void mcmc_dump_close(mcmc * m) {
	unsigned int i;
	if (m->files == NULL)
		return;
	for (i = 0; i < get_n_par(m); i++)
		if (m->files[i] != NULL)
			fclose(m->files[i]);
	free(m->files);
	m->files = NULL;
}
