functions{
	// transform node heights to proportion, except for the root
	real[] transform(real[] p, real rootHeight, int[,] map, real[] lowers){
		int S = size(p)+2;
		int nodeCount = S*2-1;
		
		real heights[S-1];
		int j = 1;
		
		heights[map[1,1]-S] = rootHeight;
		for( i in 2:nodeCount ){
			// internal node: transform
			if(map[i,1] > S){
				heights[map[i,1]-S] = lowers[map[i,1]] + (heights[map[i,2]-S] - lowers[map[i,1]])*p[j];
				j += 1;
			}
		}
		return heights;
	}
	

	real oneOnX_log(real x){
		return -log(x);
	}


	real constant_coalescent_log(real[] heights, real popSize, int[,] map, real[] lowers){
		int S = size(heights)+1; // number of leaves from the number of internal nodes
		int nodeCount = size(heights) + S;

		real logP = 0.0;
		real lineageCount = 0.0; // first 2 intervals are sampling events

		int indices[nodeCount];
		int childCounts[nodeCount];
		real times[nodeCount];

		real start;
		real finish;
		real interval;
		real logPopSize = log(popSize);

		for( i in 1:nodeCount ){
			// internal node: transform
			if(map[i,1] > S){
				times[map[i,1]] = heights[map[i,1]-S];
				childCounts[map[i,1]] = 2;
			}
			else{
				times[map[i,1]] = lowers[map[i,1]];
				childCounts[map[i,1]] = 0;
			}
		}

		// calculate intervals
		indices = sort_indices_asc(times);

		// first tip
		start = times[indices[1]];

		for (i in 1:nodeCount) {
			finish = times[indices[i]];
			
			interval = finish - start;
			if(interval != 0.0){
				logP -= interval*((lineageCount*(lineageCount-1.0))/2.0)/popSize;
			}
			
			// sampling event
			if (childCounts[indices[i]] == 0) {
				lineageCount += 1.0;
			}
			// coalescent event
			else {
				lineageCount -= 1.0;
				logP -= logPopSize;
			}
			
			start = finish;
		}

		return logP;
	}
	
	real phylo_loglik(vector blens, vector subst, vector freqs, vector rs, vector ps);

}

data{
	int <lower=0> L;                      // alignment length
	int <lower=0> S;                      // number of tips
	real<lower=0,upper=1> tipdata[S,L,4]; // alignment as partials
	int <lower=0,upper=2*S> peel[S-1,3];  // list of nodes for peeling
	real weights[L];
	int map[2*S-1,2];                     // list of node in preorder [node,parent]
	int C;
	real lower_root;
	real lowers[2*S-1]; // list of lower bounds for each internal node (for reparametrization)
	vector<lower=0>[4] frequencies_alpha; // parameters of the prior on frequencies
}

transformed data{
	int bcount = 2*S-2; // number of branches
	int nodeCount = 2*S-1; // number of nodes
	int pCount = S-2; // number of proportions
}

parameters{
	real<lower=0.1> wshape;
	real <lower=0,upper=1> props[pCount]; // proportions
	real <lower=0> rate;
	real <lower=lower_root> height; // root height
	real <lower=0> theta;
	real<lower=0> kappa;
	simplex[4] freqs;
}

transformed parameters{
	vector[C] ps = rep_vector(1.0/C, C);
	vector[C] rs;
	real <lower=0> heights[S-1];

	
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

	heights = transform(props, height, map, lowers);
}

model{
	vector [bcount] blens; // branch lengths

	wshape ~ exponential(1.0);
	rate ~ exponential(1000);
	theta ~ oneOnX();
	heights ~ constant_coalescent(theta, map, lowers);
	kappa ~ lognormal(1.0,1.25);
	freqs ~ dirichlet(frequencies_alpha);

	
	// populate blens from heights array in preorder
	for( j in 2:nodeCount ){
		// internal node
		if(map[j,1] > S){
			blens[map[j,1]] = rate*(heights[map[j,2]-S] - heights[map[j,1]-S]);
		}
		else{
			blens[map[j,1]] = rate*(heights[map[j,2]-S] - lowers[map[j,1]]);
		}
	}

		target += phylo_loglik(blens, rep_vector(kappa, 1), freqs, rs, ps);

	
	// add log det jacobian
	for( i in 2:nodeCount ){
		// skip leaves
		if(map[i,1] > S ){
			target += log(heights[map[i,2]-S] - lowers[map[i,1]]);
		}
	}

}

