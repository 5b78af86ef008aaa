/*
 * apm_analyse.c -- the `analyse` phase: marginal distributions and the model evidence.
 *
 * Both work from the files `run` wrote, exactly like the reference (src/analyse.c), so
 * runs made with either program can be analysed with either program.  In addition, when
 * `run` left its on-device accumulators in `run_statistics`, the evidence is also printed
 * from those: the prob-chain dumps carry only 7 significant digits ("%6e", reference
 * src/parallel_tempering.c:399), the accumulators are full fp64 sums.
 */
#include <gsl/gsl_histogram.h>
#include <gsl/gsl_sf.h>
#include "apm_session.h"

/* chain structs of one ensemble without an engine: analyse never touches the GPU */
static mcmc ** chains_from_files(void) {
	mcmc ** chains = setup_chains();
	read_calibration_file(chains, N_BETA);
	return chains;
}

static void free_chains(mcmc ** chains) {
	int i;
	for (i = N_BETA - 1; i >= 0; i--) {
		free(chains[i]->additional_data);
		if (i != 0)
			set_data(chains[i], NULL);
		free(mcmc_free(chains[i]));
	}
	free(chains);
}

static void print_evidence(double data_logprob) {
	/* reference src/analyse.c:95-112 */
	const double l3 = gsl_sf_log(3), l10 = gsl_sf_log(10), l30 = gsl_sf_log(30), l100 = gsl_sf_log(100);
	printf("Model probability ln(p(D|M, I)): [about 10^%.0f] %.5f"
			"\n"
			"\nTable to compare support against other models (Jeffrey):\n"
			" other model ln(p(D|M,I)) | supporting evidence for this model\n"
			" --------------------------------- \n"
			"        >  %04.1f \tnegative (supports other model)\n"
			"  %04.1f .. %04.1f \tBarely worth mentioning\n"
			"  %04.1f .. %04.1f \tSubstantial\n"
			"  %04.1f .. %04.1f \tStrong\n"
			"  %04.1f .. %04.1f \tVery strong\n"
			"        <  %04.1f \tDecisive\n", data_logprob / l10, data_logprob, data_logprob, data_logprob,
			data_logprob - l3, data_logprob - l3, data_logprob - l10, data_logprob - l10, data_logprob - l30,
			data_logprob - l30, data_logprob - l100, data_logprob - l100);
	printf("\nbe careful.\n");
}

/* thermodynamic integration by the reference's right-endpoint rectangle rule
 * (src/analyse.c:82-93): sums[k] = mean of (prob - prior) of chain k divided by beta_k */
static double integrate_over_beta(mcmc ** chains, const double * sums, unsigned int n_beta) {
	double previous_beta = 0, total = 0;
	unsigned int j;
	for (j = n_beta - 1;; j--) {
		assert(get_beta(chains[j]) > previous_beta);
		total += sums[j] * (get_beta(chains[j]) - previous_beta);
		if (j == 0)
			break;
		previous_beta = get_beta(chains[j]);
	}
	return total;
}

static int evidence_from_dumps(mcmc ** chains, unsigned int n_beta, double * result) {
	/* the files are parsed side by side, one thread per file (fscanf is what `analyse` spends
	 * its time in); every file's sum is accumulated in file order, like the reference's */
	double * sums = (double *) calloc(n_beta, sizeof(double));
	int * status = (int *) calloc(n_beta, sizeof(int)); /* 1 = not found, 2 = empty */
	int i, failed = 0;
	for (i = 0; i < (int) n_beta; i++)
		printf("reading probabilities of chain %d\r", i);
	fflush(stdout);
#pragma omp parallel for schedule(dynamic, 1)
	for (i = 0; i < (int) n_beta; i++) {
		unsigned long n = 0;
		double w, v, sum = 0;
		char name[64];
		FILE * f;
		snprintf(name, sizeof(name), "prob-chain%d.dump", i);
		f = fopen(apm_out_path(name), "r");
		if (f == NULL) {
			status[i] = 1;
			continue;
		}
		while (!feof(f)) {
			if (fscanf(f, "%le\t%le", &w, &v) == 2) {
				sum += v;
				n++;
			} else if (!feof(f)) {
				int c = fgetc(f); /* skip what cannot be parsed */
				(void) c;
			}
		}
		fclose(f);
		if (n == 0) {
			status[i] = 2;
			continue;
		}
		sums[i] = sum / get_beta(chains[i]) / n;
	}
	for (i = 0; i < (int) n_beta && !failed; i++) {
		if (status[i] == 1)
			fprintf(stderr, "calculating data probability failed: file prob-chain%d.dump not found\n", i);
		else if (status[i] == 2)
			fprintf(stderr, "calculating data probability failed: no data points found in prob-chain%d.dump\n", i);
		failed = status[i] != 0;
	}
	if (!failed)
		*result = integrate_over_beta(chains, sums, n_beta);
	free(sums);
	free(status);
	return failed;
}

static int evidence_from_accumulators(mcmc ** chains, unsigned int n_beta, double * result) {
	FILE * f = fopen(apm_out_path("run_statistics"), "r");
	double * sums;
	unsigned int i, j, n_par = get_n_par(chains[0]);
	int ok = 1;
	if (f == NULL)
		return 1;
	sums = (double *) calloc(n_beta, sizeof(double));
	for (i = 0; i < n_beta && ok; i++) {
		int k;
		unsigned long long n;
		double beta, mean_dl, skip;
		ok = fscanf(f, "%d %lf %llu %lf", &k, &beta, &n, &mean_dl) == 4 && k == (int) i && n > 0;
		for (j = 0; j < 2 * n_par && ok; j++)
			ok = fscanf(f, "%lf", &skip) == 1;
		if (ok)
			sums[i] = mean_dl / get_beta(chains[i]);
	}
	fclose(f);
	if (ok)
		*result = integrate_over_beta(chains, sums, n_beta);
	free(sums);
	return ok ? 0 : 1;
}

static int file_exists(const char * name) {
	FILE * f = fopen(apm_out_path(name), "r");
	if (f == NULL)
		return 0;
	fclose(f);
	return 1;
}

void analyse_data_probability(void) {
	double sum = 0, sum2 = 0, acc_sum = 0;
	int e, n_ok = 0, n_acc = 0;
	for (e = 0; e < N_ENSEMBLES; e++) {
		mcmc ** chains;
		double lnz;
		apm_set_output_dir(e);
		chains = chains_from_files();
		if (!file_exists("prob-chain0.dump") && evidence_from_accumulators(chains, N_BETA, &lnz) == 0) {
			/* a run without text dumps (APM_NO_DUMPS): the evidence block from the accumulators */
			if (N_ENSEMBLES > 1)
				printf("ensemble %d: ", e);
			print_evidence(lnz);
			sum += lnz;
			sum2 += lnz * lnz;
			n_ok++;
		} else if (evidence_from_dumps(chains, N_BETA, &lnz) == 0) {
			if (N_ENSEMBLES > 1)
				printf("ensemble %d: ", e);
			print_evidence(lnz);
			sum += lnz;
			sum2 += lnz * lnz;
			n_ok++;
		}
		if (evidence_from_accumulators(chains, N_BETA, &lnz) == 0) {
			printf("%sModel probability from the on-device accumulators (full precision): %.5f\n",
					N_ENSEMBLES > 1 ? "            " : "", lnz);
			acc_sum += lnz;
			n_acc++;
		}
		free_chains(chains);
	}
	apm_set_output_dir(-1);
	if (N_ENSEMBLES > 1 && n_ok > 1) {
		const double mean = sum / n_ok;
		printf("\n%d independent ensembles: ln(p(D|M, I)) = %.5f +- %.5f (standard error of the mean)\n", n_ok,
				mean, sqrt((sum2 / n_ok - mean * mean) / (n_ok - 1)));
		if (n_acc == n_ok)
			printf("from the accumulators: %.5f\n", acc_sum / n_acc);
	}
}

/* ---- marginal distributions: reference src/analyse.c:115-285, src/histogram.c:34-106 ---- */
static gsl_histogram * uniform_histogram(int nbins, double min, double max) {
	gsl_histogram * h = gsl_histogram_alloc(nbins);
	gsl_histogram_set_ranges_uniform(h, min, max);
	h->range[h->n] += (max - min) / 10000; /* so that the maximum falls into the last bin */
	return h;
}

static void fill_from_file(gsl_histogram * h, const char * filename, double * min, double * max) {
	FILE * f = fopen(filename, "r");
	double v;
	if (f == NULL) {
		fprintf(stderr, "error opening file %s\n", filename);
		perror("file could not be opened");
		exit(1);
	}
	while (fscanf(f, "%lf", &v) == 1) {
		if (h != NULL)
			gsl_histogram_increment(h, v);
		if (min != NULL && v < *min)
			*min = v;
		if (max != NULL && v > *max)
			*max = v;
	}
	if (!feof(f)) {
		fprintf(stderr, "field could not be read in %s\n", filename);
		exit(1);
	}
	fclose(f);
}

/* batch-means estimate of the Monte Carlo error of the mean (reference src/analyse.c:115-142) */
static double batch_means_error(const double mean, const char * filename, unsigned long batchsize) {
	FILE * f = fopen(filename, "r");
	unsigned long n = 0;
	int nbatches = 0;
	double v, batchsum = 0, errorsum = 0;
	if (f == NULL)
		return 0;
	while (fscanf(f, "%lf", &v) == 1) {
		n++;
		batchsum += v;
		if (n % batchsize == batchsize - 1) {
			const double batchmean = batchsum / batchsize;
			errorsum += pow(batchmean - mean, 2);
			batchsum = 0;
			nbatches++;
		}
	}
	fclose(f);
	return sqrt(errorsum / nbatches);
}

/* run_marginals: the histogram bins and batch means of the recorded chains as `run` took them from
 * the on-device accumulators (format: apm_phases.c, write_run_marginals) */
typedef struct {
	int n_par, n_bins, n_slots;
	unsigned long batch;
	unsigned long long * n_values, * n_batches; /* [n_slots] */
	unsigned long long * counts;                /* [n_slots][n_par][n_bins] */
	double ** means;                            /* [n_slots * n_par] -> [n_batches] */
} run_marginals;

static void free_run_marginals(run_marginals * m) {
	int i;
	if (m == NULL)
		return;
	if (m->means != NULL)
		for (i = 0; i < m->n_slots * m->n_par; i++)
			free(m->means[i]);
	free(m->means); free(m->counts); free(m->n_values); free(m->n_batches);
	free(m);
}

static run_marginals * read_run_marginals(void) {
	FILE * f = fopen(apm_out_path("run_marginals"), "r");
	run_marginals * m;
	int k, j, b, ok;
	if (f == NULL)
		return NULL;
	m = (run_marginals *) calloc(1, sizeof(*m));
	ok = fscanf(f, " marginals %d %d %lu %d", &m->n_par, &m->n_bins, &m->batch, &m->n_slots) == 4
			&& m->n_par > 0 && m->n_bins > 0 && m->n_slots > 0;
	if (ok) {
		m->n_values = (unsigned long long *) calloc(m->n_slots, sizeof(*m->n_values));
		m->n_batches = (unsigned long long *) calloc(m->n_slots, sizeof(*m->n_batches));
		m->counts = (unsigned long long *) calloc((size_t) m->n_slots * m->n_par * m->n_bins, sizeof(*m->counts));
		m->means = (double **) calloc((size_t) m->n_slots * m->n_par, sizeof(double *));
	}
	for (k = 0; ok && k < m->n_slots; k++) {
		int chain;
		ok = fscanf(f, " chain %d %llu %llu", &chain, &m->n_values[k], &m->n_batches[k]) == 3 && chain == k;
		for (j = 0; ok && j < m->n_par; j++) {
			double * bm = (double *) calloc(m->n_batches[k] + 1, sizeof(double));
			m->means[k * m->n_par + j] = bm;
			ok = fscanf(f, " counts") == 0;
			for (b = 0; ok && b < m->n_bins; b++)
				ok = fscanf(f, "%llu", &m->counts[((size_t) k * m->n_par + j) * m->n_bins + b]) == 1;
			ok = ok && fscanf(f, " means") == 0;
			for (b = 0; ok && b < (int) m->n_batches[k]; b++)
				ok = fscanf(f, "%lf", &bm[b]) == 1;
		}
	}
	fclose(f);
	if (!ok) {
		free_run_marginals(m);
		return NULL;
	}
	return m;
}

/* calc_mcmc_error (reference src/analyse.c:115-142) with the batch means at hand */
static double batch_means_error_from(const double mean, const double * batch_mean, unsigned long long nbatches) {
	double errorsum = 0;
	unsigned long long b;
	for (b = 0; b < nbatches; b++)
		errorsum += pow(batch_mean[b] - mean, 2);
	return sqrt(errorsum / nbatches);
}

/* (called for all parameters side by side: what it has to say goes to `report`, printed in
 * parameter order by the caller) */
static void marginal_distribution(mcmc ** chains, unsigned int n_beta, unsigned int param, char * report,
		size_t report_size, const run_marginals * acc) {
	size_t used = 0;
	const char * paramname = get_params_descr(chains[0])[param];
	double lo = get_params_min_for(chains[0], param), hi = get_params_max_for(chains[0], param);
	char in_name[APM_PATH_MAX], out_name[APM_PATH_MAX];
	gsl_histogram * h;
	double iter, mean, sigma;
	unsigned int i, filecount = 1;
	int from_accumulators = 0;
	FILE * out;
#ifdef HISTOGRAMS_ALLCHAINS
	filecount = n_beta;
#else
	(void) n_beta;
#endif
#ifdef HISTOGRAMS_MINMAX
	{
		double dmin = HUGE_VAL, dmax = -HUGE_VAL;
		for (i = 0; i < filecount; i++) {
			snprintf(in_name, sizeof(in_name), "%s-chain-%d.prob.dump", paramname, i);
			fill_from_file(NULL, apm_out_path(in_name), &dmin, &dmax);
		}
		lo = dmin;
		hi = dmax;
	}
#endif
	h = uniform_histogram(NBINS, lo, hi);
	/* the parameter dumps are there: read them, like the reference.  They are not (a run with
	 * APM_NO_DUMPS): the same bins as `run` counted them on the device. */
	snprintf(in_name, sizeof(in_name), "%s-chain-%d.prob.dump", paramname, 0);
	if (acc != NULL && !file_exists(in_name) && acc->n_bins == NBINS && acc->n_slots >= (int) filecount
			&& (int) param < acc->n_par) {
		unsigned int b;
		for (i = 0; i < filecount; i++)
			for (b = 0; b < NBINS; b++)
				h->bin[b] += (double) acc->counts[((size_t) i * acc->n_par + param) * NBINS + b];
		from_accumulators = 1;
	} else {
		for (i = 0; i < filecount; i++) {
			snprintf(in_name, sizeof(in_name), "%s-chain-%d.prob.dump", paramname, i);
			fill_from_file(h, apm_out_path(in_name), NULL, NULL);
		}
	}
	iter = gsl_histogram_sum(h);
	gsl_histogram_scale(h, (hi - lo) / NBINS / iter);
	snprintf(out_name, sizeof(out_name), "%s.histogram", paramname);
	out = fopen(apm_out_path(out_name), "w");
	assert(out != NULL);
	gsl_histogram_fprintf(out, h, DUMP_FORMAT, DUMP_FORMAT);
	fclose(out);
	mean = gsl_histogram_mean(h);
	sigma = gsl_histogram_sigma(h);
	for (i = 0; i < filecount; i++) {
		double err;
		snprintf(in_name, sizeof(in_name), "%s-chain-%d.prob.dump", paramname, i);
		if (from_accumulators)
			err = batch_means_error_from(mean, acc->means[(size_t) i * acc->n_par + param], acc->n_batches[i]);
		else
			err = batch_means_error(mean, apm_out_path(in_name), (unsigned long) sqrt(iter));
		if (used < report_size)
			used += snprintf(report + used, report_size - used, "mcmc error estimate of %s: %f %s\n", paramname, err,
					err > sigma * 0.01 ? "** high!" : " (ok)");
	}
	if (used < report_size)
		snprintf(report + used, report_size - used, "Note: Include a error estimate in your publication!\n");
	gsl_histogram_free(h);
}

void analyse_marginal_distributions(void) {
	int e;
	for (e = 0; e < N_ENSEMBLES; e++) {
		mcmc ** chains;
		unsigned int i, n_par;
		FILE * plot;
		apm_set_output_dir(e);
		chains = chains_from_files();
		n_par = get_n_par(chains[0]);
		{
			enum { REPORT = 256 * (N_BETA + 2) };
			char * reports = (char *) calloc(n_par, REPORT);
			run_marginals * acc = read_run_marginals();
			int p;
			for (i = 0; i < n_par; i++)
				printf("reading values: chain %3d parameter %s   \r", 0, get_params_descr(chains[0])[i]);
			fflush(stdout);
#pragma omp parallel for schedule(dynamic, 1)
			for (p = 0; p < (int) n_par; p++)
				marginal_distribution(chains, N_BETA, (unsigned int) p, reports + (size_t) p * REPORT, REPORT, acc);
			for (i = 0; i < n_par; i++)
				fputs(reports + (size_t) i * REPORT, stdout);
			free(reports);
			free_run_marginals(acc);
		}
		plot = fopen(apm_out_path("marginal_distributions.gnuplot"), "w");
		assert(plot != NULL);
		fprintf(plot, "# set terminal png size %d,%d; set output \"marginal_distributions.png\"\n", 600,
				300 * n_par);
		fprintf(plot, "set multiplot\n");
		fprintf(plot, "set size 1,%f\n", 1. / n_par);
		for (i = 0; i < n_par; i++) {
			fprintf(plot, "set origin 0,%f\n", (n_par - i - 1) * 1. / n_par);
			fprintf(plot, "plot \"%s.histogram\" u 1:3 title \"%s\" " GNUPLOT_STYLE "\n",
					get_params_descr(chains[0])[i], get_params_descr(chains[0])[i]);
		}
		fprintf(plot, "unset multiplot\n");
		fclose(plot);
		free_chains(chains);
	}
	apm_set_output_dir(-1);
}
