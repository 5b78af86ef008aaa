/*
 * apm_main.c -- the phase front-end of <model>.exe (mirrors reference
 * apps/generic_main.c:112-159: same phases, same argument forms).
 */
#include "apm_session.h"

#define STR_(x) #x
#define STR(x) STR_(x)

static const char * progname = "apemost";

static void usage(void) {
	fprintf(stderr, "SYNOPSIS: %s <phase> <...>\n\n", progname);
	fprintf(stderr,
			"\t-h, help\tthis text\n"
			"\tphase is one of: \n"
			"\t\tcheck          \toutput which parameters and files will be used,\n"
			"\t\t               \tcheck that they are there and that the GPU model agrees\n"
			"\t\t               \twith the host calc_model (if one is linked)\n"
			"\t\tcalibrate_first\tcalibrate first chain (beta = 1)\n"
			"\t\tcalibrate_rest \tcalibrate remaining chains (beta < 1)\n"
			"\t\trun [--append] \tcreate and dump sampling data\n"
			"\t\t               \twithout adding --append, existing data is overwritten\n"
			"\t\tanalyse        \tmarginal distributions and the data probability\n"
			"\t\thelp <phase>   \tprint more information about a phase\n\n"
			"The working directory must hold the files '" PARAMS_FILENAME "' and '" DATA_FILENAME "'.\n"
			"Algorithm parameters are compile-time macros (CCFLAGS=\"-DN_BETA=12 ...\" make <model>.exe);\n"
			"`%s check` prints the values this binary was built with.\n"
			"Environment: GSL_RNG_SEED (seed of the counter RNG), APM_DEVICE (CUDA device ordinal).\n",
			progname);
}

static void help_phase(const char * phase) {
	if (strcmp(phase, "check") == 0)
		printf("Phase 'check'\n\nPrerequisites:\n\tnone\nDoes:\n\tChecks that " PARAMS_FILENAME " and "
				DATA_FILENAME " are readable, prints the compiled configuration,\n\tevaluates the model at"
				" the start values on the GPU (and on the host if calc_model is linked).\n");
	else if (strcmp(phase, "calibrate_first") == 0)
		printf("Phase 'calibrate_first'\n\nPrerequisites:\n\tparameters file " PARAMS_FILENAME "\n\tdata file "
				DATA_FILENAME "\nProvides:\n\t" CALIBRATION_FILE " (first line), " PARAMS_FILENAME
				"_suggested, calibration_progress.data\nDoes:\n\tburn-in and step width calibration of the"
				" first chain (beta = 1)\n");
	else if (strcmp(phase, "calibrate_rest") == 0)
		printf("Phase 'calibrate_rest'\n\nPrerequisites:\n\tfirst line of " CALIBRATION_FILE
				"\nProvides:\n\t" CALIBRATION_FILE " (all chains), calibration_summary\nDoes:\n\tchooses the"
				" beta ladder, predicts and calibrates the step widths of the remaining chains\n");
	else if (strcmp(phase, "run") == 0)
		printf("Phase 'run'\n\nPrerequisites:\n\t" CALIBRATION_FILE "\nProvides:\n\t<parameter>-chain-0.prob.dump,"
				" prob-chain<k>.dump, acceptance_rate.dump, run_statistics\nDoes:\n\tparallel tempering sampling"
				" until MAX_ITERATIONS or Ctrl-C; SIGUSR1 flushes the dumps\n");
	else if (strcmp(phase, "analyse") == 0)
		printf("Phase 'analyse'\n\nPrerequisites:\n\tthe dump files of 'run'\nProvides:\n\t<parameter>.histogram,"
				" marginal_distributions.gnuplot, the model probability on stdout\n");
	else
		usage();
}

int main(int argc, char ** argv) {
	progname = argv[0];
	if (argc < 2) {
		fprintf(stderr, "No phase specified.\n");
		usage();
		return 0;
	}
	if (strcmp(argv[1], "help") == 0 || strcmp(argv[1], "-h") == 0) {
		if (argc == 3)
			help_phase(argv[2]);
		else
			usage();
	} else if (strcmp(argv[1], "check") == 0) {
		check();
	} else if (strcmp(argv[1], "calibrate_first") == 0) {
		calibrate_first();
	} else if (strcmp(argv[1], "calibrate_rest") == 0) {
		calibrate_rest();
	} else if (strcmp(argv[1], "run") == 0) {
		if (argc == 3 && strcmp(argv[2], "--append") == 0)
			prepare_and_run_sampler(MAX_ITERATIONS, 1);
		else if (argc == 2)
			prepare_and_run_sampler(MAX_ITERATIONS, 0);
		else {
			fprintf(stderr, "You are doing it wrong.\nDid you want to write --append?\n");
			usage();
		}
	} else if (strcmp(argv[1], "analyse") == 0) {
		if (argc == 3 && (strcmp(argv[2], "marginal") == 0 || strcmp(argv[2], "model") == 0)) {
			analyse_marginal_distributions(); /* both spellings do this in the reference too */
		} else if (argc == 2) {
			analyse_marginal_distributions();
			analyse_data_probability();
		} else {
			fprintf(stderr, "You are doing it wrong.\nDid you want to write 'analyse marginal' or 'analyse model'?\n");
			usage();
		}
	} else {
		fprintf(stderr, "You are doing it wrong.\n");
		usage();
	}
	return 0;
}

/* ------------------------------------------------------------------ check */
static void check_file(const char * path) {
	FILE * f = fopen(path, "r");
	printf("\t%s: %s\n", path, f != NULL ? "found" : "NOT FOUND");
	if (f != NULL)
		fclose(f);
}

void check(void) {
	printf("%s: Checking environment:\n\nFiles:\n", progname);
	check_file(PARAMS_FILENAME);
	check_file(DATA_FILENAME);
	printf("\nModel:\n\tdevice model: %s (id %d)\n", APM_MODEL_NAME, (int) APM_MODEL_ID);
#ifdef APM_HAVE_HOST_MODEL
	printf("\thost calc_model: linked\n");
#else
	printf("\thost calc_model: not linked\n");
#endif
	printf("\nFine-tuning the algorithm:\n");
	printf("\tBETA_ALIGNMENT: %s\n", STR(BETA_ALIGNMENT));
	printf("\tN_ENSEMBLES: %d\n", (int) N_ENSEMBLES);
	printf("\tN_BETA: %d\n", (int) N_BETA);
	printf("\tBETA_0: %f\n", (double) BETA_0);
	printf("\tBURN_IN_ITERATIONS: %d\n", (int) BURN_IN_ITERATIONS);
	printf("\tTARGET_ACCEPTANCE_RATE: %f\n", (double) TARGET_ACCEPTANCE_RATE);
	printf("\tMAX_AR_DEVIATION: %f\n", (double) MAX_AR_DEVIATION);
	printf("\tITER_LIMIT: %d\n", (int) ITER_LIMIT);
	printf("\tMUL: %f\n", (double) MUL);
	printf("\tN_SWAP: %d\n", (int) N_SWAP);
	printf("\tCIRCULAR_PARAMS: %s\n", STR((CIRCULAR_PARAMS)));
#ifdef SKIP_CALIBRATE_ALLCHAINS
	printf("\tSKIP_CALIBRATE_ALLCHAINS: enabled (calibrating only 2 chains)\n");
#else
	printf("\tSKIP_CALIBRATE_ALLCHAINS: disabled (calibrating all chains)\n");
#endif
#if defined(PROPOSAL_UNIFORM)
	printf("\tPROPOSAL: uniform proposal distribution\n");
#elif defined(PROPOSAL_LOGISTIC)
	printf("\tPROPOSAL: logistic proposal distribution\n");
#else
	printf("\tPROPOSAL: gaussian/normal proposal distribution\n");
#endif
	printf("\tRANDOMSWAP: Random swapping: ");
#ifdef RANDOMSWAP
	printf("on\n");
#else
	printf("off\n");
#endif
#if defined(CALIBRATE_MULTILIN)
	printf("\tCALIBRATE_MULTILIN: on (assess_acceptance_rate + markov_chain_calibrate_multilinear_regression)\n");
#elif defined(CALIBRATE_ALTERNATE)
	printf("\tCALIBRATE_ALTERNATE: on (assess_acceptance_rate + markov_chain_calibrate_alt)\n");
#else
	printf("\tCALIBRATE_ALTERNATE: off (markov_chain_calibrate_orig)\n");
#endif
#ifdef ADAPT
	printf("\tADAPT: on (1%% step width rescaling per round once 20000 moves are counted)\n");
#else
	printf("\tADAPT: off\n");
#endif
#ifdef APM_EXACT_SWAP
	printf("\tAPM_EXACT_SWAP: swaps exchange the likelihood with the position (not the reference's behaviour)\n");
#else
	printf("\tAPM_EXACT_SWAP: off (swap and revert behave exactly like the reference)\n");
#endif
	printf("\nRunning:\n");
	if (MAX_ITERATIONS != 0)
		printf("\tMAX_ITERATIONS: Stops after %lu iterations\n", (unsigned long) MAX_ITERATIONS);
	else
		printf("\tMAX_ITERATIONS: Run indefinitely long\n");
	printf("\tPRINT_PROB_INTERVAL: %d\n", (int) PRINT_PROB_INTERVAL);
#ifdef DUMP_ALL_CHAINS
	printf("\tDUMP_ALL_CHAINS: on\n");
#else
	printf("\tDUMP_ALL_CHAINS: off (parameter dumps of chain 0 only)\n");
#endif
	{
		/* the model at the start values: device, and host if calc_model is linked */
		FILE * p = fopen(PARAMS_FILENAME, "r"), *d = fopen(DATA_FILENAME, "r");
		if (p != NULL && d != NULL) {
			apm_session * s;
			int zero = 0;
			fclose(p);
			fclose(d);
			s = apm_session_open();
			apm_session_calc_model(s, &zero, 1);
			printf("\nModel at the start values:\n\tdevice: prob = " DUMP_FORMAT "  prior = " DUMP_FORMAT "\n",
					get_prob(s->chains[0]), get_prior(s->chains[0]));
#ifdef APM_HAVE_HOST_MODEL
			{
				const double dev_prob = get_prob(s->chains[0]);
				calc_model(s->chains[0], NULL);
				printf("\thost:   prob = " DUMP_FORMAT "  prior = " DUMP_FORMAT "\n", get_prob(s->chains[0]),
						get_prior(s->chains[0]));
				printf("\trelative difference: %.3e\n", fabs(dev_prob - get_prob(s->chains[0]))
						/ (fabs(get_prob(s->chains[0])) > 0 ? fabs(get_prob(s->chains[0])) : 1));
			}
#endif
			apm_session_close(s);
		} else {
			if (p != NULL)
				fclose(p);
			if (d != NULL)
				fclose(d);
		}
	}
}
