functions{	real phylo_loglik(vector blens, vector subst, vector freqs, vector rs, vector ps);

}

data{
	int <lower=0> L;                      // alignment length
	int <lower=0> S;                      // number of tips
	real<lower=0,upper=1> tipdata[S,L,4]; // alignment as partials
	int <lower=0,upper=2*S> peel[S-1,3];  // list of nodes for peeling
	real weights[L];
	int C;
	vector<lower=0>[4] frequencies_alpha; // parameters of the prior on frequencies
	vector<lower=0>[6] rates_alpha;       // parameters of the prior on rates
}

transformed data{
	int bcount = 2*S-3; // number of branches
	int nodeCount = 2*S-1; // number of nodes
}

parameters{
	real<lower=0.1> wshape;
	vector<lower=0> [bcount] blens; // branch lengths
	simplex[6] rates;
	simplex[4] freqs;
}

transformed parameters{
	vector[C] ps = rep_vector(1.0/C, C);
	vector[C] rs;

	
		{
			real m = 0;
			for(i in 1:C){
				rs[i] = pow(-log(1.0 - (2.0*(i-1)+1.0)/(2.0*C)), 1.0/wshape);
			}
			m = sum(rs)/C;
			for(i in 1:C){
				rs[i] /= m;		
			}
		}

}

model{

	wshape ~ exponential(1.0);
	blens ~ exponential(10);
	rates ~ dirichlet(rates_alpha);
	freqs ~ dirichlet(frequencies_alpha);

		target += phylo_loglik(blens, rates, freqs, rs, ps);

}

